"""Drop-in test of the host adapter's answer to inputs the device record scanner rejects (-m gpu).

The scanner takes FASTA (single- or multi-line) and 4-line FASTQ.  BankFasta also accepts multi-line FASTQ (quality lines
are consumed until their total length reaches the read length, G/src/gatb/bank/impl/BankFasta.cpp:542-553): the scanner
answers DSKGPU_ERR_FORMAT, host/GpuSortingCount.hpp resets the context and feeds the same banks through the reference's
own parser (IBank::iterator -> dskgpu_push_reads).  The output must be what the reference `dsk` writes for the same file.
(Checked on B200 by tools/run_cli_check.sh before this test was written; kept in its own file, last in the suite.)"""
import os

import pytest

from util import INPUTS
import test_cli_dropin as cli

pytestmark = pytest.mark.gpu


def multiline_fastq(src, dst):
    """every sequence line and every quality line of a 4-line FASTQ split in two"""
    out = []
    for i, ln in enumerate(open(src, "rb").read().split(b"\n")):
        if i % 4 in (1, 3) and ln:
            h = len(ln) // 2
            out += [ln[:h], ln[h:]]
        else:
            out.append(ln)
    open(dst, "wb").write(b"\n".join(out))


@cli.need_bins
@pytest.mark.skipif(not os.path.exists(os.path.join(cli.REFBIN, "dsk")), reason="reference dsk binary not on this box")
def test_cli_multiline_fastq_goes_through_the_reference_parser(tmp_path):
    tmp = str(tmp_path)
    fq = os.path.join(tmp, "ml.fastq")
    multiline_fastq(os.path.join(INPUTS, "reads.fastq"), fq)
    a, b = os.path.join(tmp, "gpu_out"), os.path.join(tmp, "ref_out")
    args = ["-file", fq, "-kmer-size", "21", "-abundance-min", "1", "-histo", "1", "-verbose", "0"]
    cli.run([cli.DSK_GPU] + args + ["-out", a], tmp)
    cli.run([os.path.join(cli.REFBIN, "dsk")] + args + ["-out", b, "-out-tmp", tmp, "-nb-cores", "2"], tmp)
    la, ha, _ = cli.read_back(a + ".h5", tmp)
    lb, hb, _ = cli.read_back(b + ".h5", tmp)
    assert len(lb) > 10000 and la == lb                      # dsk2ascii | sort
    assert ha == hb                                          # gatb-h5dump -y -d histogram/histogram
    assert open(a + ".histo", "rb").read() == open(b + ".histo", "rb").read()
