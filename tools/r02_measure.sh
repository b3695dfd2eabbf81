set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
python tools/run_configs.py > gpurun_out/r02_configs.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_ncu_l.log 2>&1
python tools/profile_loop.py 1024 60 > gpurun_out/r02_pl_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tw_closed_loop -c 1 -s 1 -o gpurun_out/r02_closed_loop -f python tools/profile_loop.py 1024 60 > gpurun_out/r02_ncu_f.log 2>&1
tail -3 gpurun_out/r02_configs.txt
