#!/usr/bin/env bash
# Runs under gpurun on 1 GPU, no python (seconds): the C++ host side `dsk_gpu` against the reference readers / binary.
#  1. BASELINE configs[0] (repo test file, k=31, abundance-min 2, -histo): md5 of the sorted dsk2ascii dump + .histo
#     against the values SURVEY.md 8(c) recorded from the reference (9905e889... / 799dd0e1...)
#  2. a multi-line FASTQ (rejected by the device scanner): the adapter must feed it through the reference's parser and
#     produce what the reference `dsk` produces
set -u
B="$PWD"; I="$B/tests/golden/inputs"; OUT="$B/gpurun_out/${1:-r01z_cli}"; mkdir -p "$OUT"; cd "$OUT"
R="$B/oracle/_ref/bin"
"$B/host/_build/dsk_gpu" -file "$I/read50x_ref10K_e001.fasta.gz" -kmer-size 31 -abundance-min 2 -out c1 -histo 1 -verbose 0 > c1.log 2>&1; echo "dsk_gpu c1 exit $?"
"$R/dsk2ascii" -file c1.h5 -out c1.txt > /dev/null 2>&1; LC_ALL=C sort c1.txt | md5sum | cut -c1-32 > c1.md5; md5sum < c1.histo | cut -c1-32 >> c1.md5; cat c1.md5
awk 'NR%4==2||NR%4==0{h=int(length($0)/2); print substr($0,1,h); print substr($0,h+1); next} {print}' "$I/reads.fastq" > ml.fastq
"$B/host/_build/dsk_gpu" -file ml.fastq -kmer-size 21 -abundance-min 1 -out ml_gpu -histo 1 -verbose 0 > ml_gpu.log 2>&1; echo "dsk_gpu multi-line fastq exit $?"
"$R/dsk" -file ml.fastq -kmer-size 21 -abundance-min 1 -out ml_ref -histo 1 -verbose 0 -nb-cores 2 > ml_ref.log 2>&1; echo "ref dsk exit $?"
"$R/dsk2ascii" -file ml_gpu.h5 -out ml_gpu.txt > /dev/null 2>&1; "$R/dsk2ascii" -file ml_ref.h5 -out ml_ref.txt > /dev/null 2>&1
LC_ALL=C sort ml_gpu.txt | md5sum; LC_ALL=C sort ml_ref.txt | md5sum; wc -l ml_gpu.txt ml_ref.txt | head -2
cmp ml_gpu.histo ml_ref.histo && echo "histo identical"
tail -3 ml_gpu.log
rm -f *.h5 ml.fastq *.txt
