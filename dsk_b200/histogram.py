"""Host side of `-abundance-min auto`: the cutoff heuristic and the threshold update between the two count passes.

Follows gatb-core's Histogram::compute_threshold (G/src/gatb/tools/misc/impl/Histogram.cpp:61-190), the cutoff
processor that feeds it (G/src/gatb/kmer/impl/CountProcessorCutoff.hpp:86-124) and
CountProcessorSolidityInfo::setAbundanceMin (G/src/gatb/kmer/impl/CountProcessorSolidity.hpp:45-66).  The only
floating-point code of the counting path: doubles, evaluated in the reference's order, truncated to u64 like its casts.
"""

MIN_AUTO_THRESHOLD = 3            # CountProcessorCutoff.hpp:88
HISTO_LENGTH = 10000


def compute_threshold(hist, min_auto_threshold=MIN_AUTO_THRESHOLD, length=HISTO_LENGTH):
    """hist[i] = number of distinct k-mers of abundance i (length+1 entries).  Returns (cutoff, nbsolids, first_peak)."""
    h = [int(x) for x in hist[:length + 1]]
    sm = [0] * (length + 1)
    sum_allk = 0
    if length >= 2:
        sm[1] = int(0.6 * float(h[1]) + 0.4 * float(h[2]))
        sum_allk += h[1]
    first_inc, idx_max, max_val = -1, -1, 0
    for i in range(2, length):
        sum_allk += h[i] * i
        sm[i] = int(0.2 * float(h[i - 1]) + 0.6 * float(h[i]) + 0.2 * float(h[i + 1]))
        if first_inc == -1 and sm[i - 1] < sm[i]:
            first_inc = i - 1
        if first_inc > 0 and sm[i] > max_val:
            max_val, idx_max = sm[i], i
    sum_allk += h[length] * length
    if first_inc == -1:
        return min_auto_threshold, 0, 0                      # early return: _nbsolids stays 0 (Histogram.cpp:101-105)
    first_peak = idx_max
    cutoff, min_val = 0, 10000000000
    for i in range(first_inc, idx_max + 1):
        if sm[i] < min_val:
            min_val, cutoff = sm[i], i
    sum_elim, max_cutoff = 0, 0
    for i in range(length + 1):
        sum_elim += h[i] * i
        if sum_allk and float(sum_elim) / float(sum_allk) >= 0.25:
            max_cutoff = i + 1
            break
    cutoff = min(cutoff, max_cutoff)
    cutoff = max(cutoff, min_auto_threshold)
    return cutoff, sum(h[cutoff:length + 1]), first_peak


def auto_thresholds(abundance_min, cutoffs):
    """setAbundanceMin: entries equal to -1 ("auto") take their bank's cutoff; banks beyond the cutoffs copy the last one."""
    if len(cutoffs) > len(abundance_min):
        raise ValueError("Unable to set abundance min values (%d values for %d banks)" % (len(cutoffs), len(abundance_min)))
    out = [c if a == -1 else a for a, c in zip(abundance_min, cutoffs)]
    return out + [out[-1]] * (len(abundance_min) - len(out))
