ncu --set full --clock-control none --import-source on -k regex:"loss_grad" -s 3 -c 1 -o gpurun_out/prof_r1b_lossgrad python tools/kbench.py --iters 1 --only multiloss_grad > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"resample_confusion" -s 3 -c 1 -o gpurun_out/prof_r1c_resample python tools/kbench.py --iters 1 --only resample > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
