set -x
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_gpu_tests_1gpu.log
python bench.py > gpurun_out/r2_bench_amos98.json 2> gpurun_out/bench_amos98.err
python bench.py --config wide > gpurun_out/r2_bench_wide.json 2> gpurun_out/bench_wide.err
for b in 2 4; do for f in 1 0; do echo "=== batch $b DUNET_FLAT=$f DUNET_FLAT_DECONV=$f DUNET_FUSED_SPLITK_NORM=$f"; DUNET_FLAT=$f DUNET_FLAT_DECONV=$f DUNET_FUSED_SPLITK_NORM=$f timeout 200 python tools/one_window.py --batch $b --reps 5 --prof --dump 2>&1; done; done > gpurun_out/r2_deep_levels_per_launch.txt
bash tools/deep_timeline_all.sh > gpurun_out/r2_deep_conv_cta_timeline_flat.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_batch4_window.csv python tools/one_window.py --batch 4 --reps 1 > gpurun_out/ncu_ll.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv3d_flat --launch-skip 8 --launch-count 13 -o gpurun_out/r2_flat_b2 -f python tools/one_window.py --batch 2 --reps 1 > gpurun_out/ncu_flat.log 2>&1
echo finished
