#!/bin/bash
mkdir -p gpurun_out
B200VFX_LIB=$PWD/gst-plugin-rs_b200/lib/libb200vfx_blocked.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "all_colors or variants or formats_strides or policy" 2>&1 | tail -3
echo "--- linear layout"; python scripts/sweep_memo.py 2>/dev/null | head -5
echo "--- blocked layout"; B200VFX_LIB=$PWD/gst-plugin-rs_b200/lib/libb200vfx_blocked.so python scripts/sweep_memo.py 2>/dev/null | head -5
echo "--- hsv memo (blocked)"; B200VFX_LIB=$PWD/gst-plugin-rs_b200/lib/libb200vfx_blocked.so python scripts/kernel_bench.py --only hsv 2>/dev/null | grep "memo" | cut -c1-200
