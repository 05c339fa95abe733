set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_final.log 2>&1; tail -3 gpurun_out/t_final.log
python bench.py > gpurun_out/bench_final_r2.log 2>&1; tail -c 300 gpurun_out/bench_final_r2.log
MGCONV_LANES=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2f.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_r2f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_wgrad_halo_pair -s 3 -c 1 -f -o gpurun_out/prof_r2f_b3g1_wgrad_pair python scratch/conv_bench.py b3g1 wgrad 2 > gpurun_out/ncu_tmp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_stem_wgrad -s 3 -c 1 -f -o gpurun_out/prof_r2f_stem_wgrad python scratch/conv_bench.py stem wgrad 2 > gpurun_out/ncu_tmp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_conv_halo_persistent -s 3 -c 1 -f -o gpurun_out/prof_r2f_b1g1_fwd_stats python scratch/conv_bench.py b1g1 fwd_stats 2 > gpurun_out/ncu_tmp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:umma_stem_kernel -s 3 -c 1 -f -o gpurun_out/prof_r2f_stem_fwd python scratch/conv_bench.py stem fwd 2 > gpurun_out/ncu_tmp.log 2>&1
ls -la gpurun_out/prof_r2f*
