#!/bin/bash
# Round profiling pass (run under gpurun): ncu launch list of a short bench step + one `--set full`
# capture per custom kernel at the BASELINE sizes (tools/kbench.py).  Reports land in gpurun_out/.
set -u
R=${1:-r2}
mkdir -p gpurun_out
# launch list of the TIMED region only (bench.py brackets it with cudaProfilerStart/Stop)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv \
    --log-file gpurun_out/launches_$R.csv python bench.py --steps 1 --warmup 2 --images 4 --no-cpu-baseline > gpurun_out/bench_under_ncu_$R.log 2>&1
tail -3 gpurun_out/bench_under_ncu_$R.log
wc -l gpurun_out/launches_$R.csv
cap() {  # name, kernel regex, kbench --only filter, launches to keep (after 3 warm-up launches each)
    ncu --set full --clock-control none --import-source on -k regex:"$2" -s 3 -c "$4" -o gpurun_out/prof_${R}_$1 \
        python tools/kbench.py --iters 1 --only "$3" > gpurun_out/ncu_$1_$R.log 2>&1
    tail -2 gpurun_out/ncu_$1_$R.log
}
cap stitch "stitch_kernel" "stitch_argmax_colour 273 tiles C9 labels" 1
cap stitch45 "stitch_kernel" "stitch_argmax_colour 45 tiles" 1
cap stitchup "stitch_up_kernel" "stitch_upsample" 1
cap resample "resample_confusion" "resample" 1
cap loss "loss_reduce|loss_grad" "multiloss_reduce B64 C9 i64,multiloss_grad" 2
cap lossfused "loss_fused" "one cooperative launch" 1
cap resize "area_resize" "fit_resize_area" 2
cap gather "gather_mask|gather_norm|gather_img" "mask_gather_encode_hist 6000x4000 S512 C9,gather_norm,tile_gather_u8 rgb" 4
cap encode "class_encode|tile_hist|tile_moments" "class_encode,profile_tiles" 4
cap netglue "upsample_concat|maxpool3x3s2|upsample_to_nchw" "upsample_concat,maxpool,upsample_nhwc_to_nchw" 3
ls -la gpurun_out | head -30
