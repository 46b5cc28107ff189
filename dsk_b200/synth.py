"""Deterministic synthetic read sets of the shapes BASELINE.json names.

genome  : i.i.d. uniform ACGT of length G                      (numpy PCG64, seed s0)
reads   : start ~ U[0, G-L], fixed length L, strand ~ Bernoulli(1/2) (reverse-complemented),
          each base substituted with probability e = round(e*65536)/65536 by a uniformly different base (seed s1)
format  : FASTA, single-line sequences, header ">r<i>"
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def genome_codes(G, seed=42):
    return np.random.default_rng(seed).integers(0, 4, G, dtype=np.uint8)


def reads_fasta(G=5_000_000, coverage=100, L=150, err=0.01, seed=42, out=None, max_reads=None, genome=None):
    """Returns (buffer, nbytes, nreads).  `out`: optional preallocated uint8 array (e.g. pinned memory)."""
    g = genome if genome is not None else genome_codes(G, seed)
    G = len(g)
    n = int(G * coverage // L)
    if max_reads is not None:
        n = min(n, max_reads)
    rng = np.random.default_rng(seed + 1)
    thr = int(round(err * 65536))
    # exact output size: per read '>r' + digits + '\n' + L + '\n'
    digits = np.ones(n, dtype=np.int64)
    lim = 10
    while lim <= max(n - 1, 1):
        digits[lim:] += 1
        lim *= 10
    total = int((digits + 2 + 1 + L + 1).sum())
    buf = out if out is not None else np.empty(total, dtype=np.uint8)
    assert buf.size >= total
    CH = 200_000
    pos = 0
    ar = np.arange(L, dtype=np.int64)
    for b in range(0, n, CH):
        e = min(n, b + CH)
        m = e - b
        starts = rng.integers(0, G - L + 1, m)
        codes = g[starts[:, None] + ar[None, :]]
        if thr:
            mask = rng.integers(0, 65536, (m, L), dtype=np.uint16) < thr
            shift = rng.integers(1, 4, (m, L), dtype=np.uint8)
            codes = np.where(mask, (codes + shift) & 3, codes)
        strand = rng.integers(0, 2, m, dtype=np.uint8).astype(bool)
        rc = (3 - codes[:, ::-1])
        codes = np.where(strand[:, None], rc, codes)
        letters = _ACGT[codes]
        # rows of equal header width are written as one 2-D block
        ids = np.arange(b, e, dtype=np.int64)
        dg = digits[b:e]
        for d in np.unique(dg):
            sel = np.nonzero(dg == d)[0]
            w = 2 + int(d) + 1 + L + 1
            blk = np.empty((len(sel), w), dtype=np.uint8)
            blk[:, 0] = ord(">")
            blk[:, 1] = ord("r")
            v = ids[sel].copy()
            for j in range(int(d) - 1, -1, -1):
                blk[:, 2 + j] = 48 + (v % 10)
                v //= 10
            blk[:, 2 + int(d)] = 10
            blk[:, 3 + int(d):3 + int(d) + L] = letters[sel]
            blk[:, -1] = 10
            # rows with the same width are contiguous in id order (ids are monotone, digits non-decreasing)
            buf[pos:pos + blk.size] = blk.reshape(-1)
            pos += blk.size
    assert pos == total
    return buf, total, n


def assembly_fasta(g, width=70, name=b"contig"):
    """the genome itself as one multi-line FASTA record (bank 0 of the -histo2D configuration)"""
    letters = _ACGT[g]
    nfull = len(g) // width
    body = np.empty((nfull, width + 1), dtype=np.uint8)
    body[:, :width] = letters[:nfull * width].reshape(nfull, width)
    body[:, width] = 10
    tail = letters[nfull * width:].tobytes()
    return b">" + name + b"\n" + body.tobytes() + (tail + b"\n" if tail else b"")


def genome_device(G, seed=42, device="cuda"):
    import torch
    gen = torch.Generator(device=device); gen.manual_seed(seed)
    return torch.randint(0, 4, (G,), dtype=torch.uint8, device=device, generator=gen)


def assembly_fasta_device(g, width=70):
    """the genome as one multi-line FASTA record, built on the device (bank 0 of the -histo2D configuration)"""
    import torch
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=g.device)
    head = torch.tensor(list(b">contig\n"), dtype=torch.uint8, device=g.device)
    nfull = g.numel() // width
    body = torch.empty((nfull, width + 1), dtype=torch.uint8, device=g.device)
    body[:, :width] = acgt[g[:nfull * width].long()].view(nfull, width)
    body[:, width] = 10
    tail = acgt[g[nfull * width:].long()]
    nl = torch.tensor([10], dtype=torch.uint8, device=g.device)
    return torch.cat([head, body.view(-1)] + ([tail, nl] if tail.numel() else []))


def reads_fasta_device(G, coverage=30, L=150, err=0.01, seed=42, device="cuda", batch=1 << 21, genome=None, nreads=None):
    """Same read model generated on the device with torch (benchmark plumbing for read sets too big to draw with numpy
    in reasonable time: the 3 Gbp-class configurations).  Fixed-width headers ">r%09d" so that every record has the same
    size and a batch is one 2-D tensor.  Returns (uint8 cuda tensor of FASTA bytes, nreads).  Not bit-identical to
    reads_fasta() (different PRNG); parity at this scale is checked through size-independent properties."""
    import torch
    gen = torch.Generator(device=device); gen.manual_seed(seed + 1)
    g = genome if genome is not None else genome_device(G, seed, device)
    G = g.numel()
    n = int(G * coverage // L) if nreads is None else int(nreads)      # nreads: this rank's slice of the read set
    W = 2 + 9 + 1 + L + 1
    out = torch.empty(n * W, dtype=torch.uint8, device=device)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    ar = torch.arange(L, device=device, dtype=torch.int64)
    pow10 = torch.tensor([10 ** (8 - j) for j in range(9)], device=device, dtype=torch.int64)
    for b in range(0, n, batch):
        m = min(batch, n - b)
        starts = torch.randint(0, G - L + 1, (m,), device=device, generator=gen)
        codes = g[starts[:, None] + ar[None, :]]
        if err > 0:
            mask = torch.rand((m, L), device=device, generator=gen) < err
            shift = torch.randint(1, 4, (m, L), dtype=torch.uint8, device=device, generator=gen)
            codes = torch.where(mask, (codes + shift) & 3, codes)
        strand = torch.randint(0, 2, (m,), device=device, generator=gen).bool()
        codes = torch.where(strand[:, None], 3 - codes.flip(1), codes)
        blk = out[b * W:(b + m) * W].view(m, W)
        blk[:, 0] = ord(">"); blk[:, 1] = ord("r")
        ids = torch.arange(b, b + m, device=device, dtype=torch.int64)
        blk[:, 2:11] = (48 + (ids[:, None] // pow10[None, :]) % 10).to(torch.uint8)
        blk[:, 11] = 10
        blk[:, 12:12 + L] = acgt[codes.long()]
        blk[:, 12 + L] = 10
        del codes, starts
    return out, n
