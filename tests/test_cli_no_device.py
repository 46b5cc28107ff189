"""CPU-side checks of the C++ host binary (host/_build/dsk_gpu): without a CUDA device the counting path must fail loudly
(there is no CPU fallback), and options outside the device path must be refused with the reference's exception type
instead of silently computing something else.  (Skipped on a box that has a GPU: there the drop-in tests run.)"""
import os
import subprocess

import pytest

from util import INPUTS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DSK_GPU = os.path.join(ROOT, "host", "_build", "dsk_gpu")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


pytestmark = [pytest.mark.skipif(not os.path.exists(DSK_GPU), reason="host/_build/dsk_gpu not built (python __graft_entry__.py)"),
              pytest.mark.skipif(has_gpu(), reason="a CUDA device is present: covered by tests/test_cli_dropin.py")]


def run_cli(tmp_path, *extra):
    cmd = [DSK_GPU, "-file", os.path.join(INPUTS, "shortread.fasta"), "-kmer-size", "15", "-abundance-min", "1",
           "-out", str(tmp_path / "out"), "-verbose", "0"] + list(extra)
    return subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True)


def test_no_device_fails_loudly_no_cpu_fallback(tmp_path):
    p = run_cli(tmp_path)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stdout + p.stderr) and "no CPU fallback" in (p.stdout + p.stderr)


@pytest.mark.parametrize("opt,val,msg", [("-histo-max", "20000", "outside 1..10000"), ("-histo-max", "0", "outside 1..10000")])
def test_options_outside_the_device_path_are_refused(tmp_path, opt, val, msg):
    p = run_cli(tmp_path, opt, val)
    assert p.returncode != 0
    assert "EXCEPTION" in (p.stdout + p.stderr) and msg in (p.stdout + p.stderr)      # refused for what it is, before any device call


@pytest.mark.parametrize("opt,val", [("-minimizer-type", "1"), ("-repartition-type", "1"), ("-histo-max", "5000")])
def test_partition_balance_options_are_accepted(tmp_path, opt, val):
    # they only move k-mers between partitions (unobservable) / shorten the histogram: accepted; without a device the run then
    # fails for the one reason it should
    p = run_cli(tmp_path, opt, val)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stdout + p.stderr) and "outside" not in (p.stdout + p.stderr)


def test_unhandled_kmer_size_message(tmp_path):
    p = subprocess.run([DSK_GPU, "-file", os.path.join(INPUTS, "shortread.fasta"), "-kmer-size", "128", "-out", str(tmp_path / "x")],
                       cwd=str(tmp_path), capture_output=True, text=True)
    assert p.returncode != 0 and "unhandled kmer size 128" in (p.stdout + p.stderr)
