timeout 900 python -u -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r01f_bench_n1.json 2> gpurun_out/r01f_bench_n1.err; tail -c 300 gpurun_out/r01f_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 60 --csv --log-file gpurun_out/r01f_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r01f_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_col2_pipe|k_row" -s 10 -c 5 -o gpurun_out/r01f_fk -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r01f_fk.log 2>&1
tail -1 gpurun_out/r01f_fk.log
