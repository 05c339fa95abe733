set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_final.log 2>&1; tail -3 gpurun_out/t_final.log
python bench.py > gpurun_out/bench_final_r2.log 2>&1; tail -c 200 gpurun_out/bench_final_r2.log
MGCONV_LANES=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2h.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_r2h.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bn_relu_pool3 -s 1 -c 1 -f -o gpurun_out/prof_r2h_stem_bn_relu_pool3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tmp.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
