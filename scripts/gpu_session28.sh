#!/bin/bash
# compute-sanitizer on the kernels added in this session (cluster/DSMEM histogram, shared-memory transposition of 3-byte
# pixels, shared-memory sub-cube, batched block sums), then ncu digests of the same kernels
mkdir -p gpurun_out/s28
K='histogram_matches_oracle or device_pointers_unaligned or rgb24_memo or blockhash_batch or memo_tile_kernel_hsvfilter'
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_colordetect.py tests/test_gpu_parity.py tests/test_gpu_variants.py -x -q -m gpu -k "$K" > gpurun_out/s28/memcheck.txt 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/s28/memcheck.txt
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_colordetect.py tests/test_gpu_parity.py tests/test_gpu_variants.py -x -q -m gpu -k "histogram_matches_oracle or rgb24_memo or memo_tile_kernel_hsvfilter" > gpurun_out/s28/racecheck.txt 2>&1
echo "racecheck rc=$?"; tail -4 gpurun_out/s28/racecheck.txt
for spec in "colordetect ramps 1" "colordetect noise 1" "colordetect ramps 10"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:colordetect_hist -s 2 -c 1 -f -o gpurun_out/s28/cd_$2_q$3 \
      python scripts/ncu_target.py --kernel colordetect --content $2 --launches 4 --quality $3 > gpurun_out/s28/ncu_cd.log 2>&1
done
ls gpurun_out/s28
