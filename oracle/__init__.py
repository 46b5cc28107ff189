"""CPU oracle for the DSK counting hot path -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs.
The product package (dsk_b200) never imports this module.
"""
from .pyoracle import (Oracle, OracleResult, count_files, histogram_threshold, kmers_of, parse_stats, mmer_lut,
                       ref_available, ref_wide_available, run_reference, kmers_of_words, read_maybe_gz, ORACLE_DIR)
