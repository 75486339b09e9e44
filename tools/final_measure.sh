set -x
python bench.py > gpurun_out/bench_own.json 2> gpurun_out/bench_own.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --layers 20 --no-cpu --no-extras > gpurun_out/launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tile_pass_lean -s 20 -c 2 -o gpurun_out/r01_lean_pass -f python tools/hea_cfg.py 30 12 c128 11:5:128 > gpurun_out/ncu_full.log 2>&1
python tools/config3_sweep.py 100 30 30,33 > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
tail -3 gpurun_out/sweep.jsonl | cut -c1-300
