mkdir -p gpurun_out
timeout 600 python bench.py --bs 64 --no-secondary --no-cpu-baseline > gpurun_out/b64.json 2> gpurun_out/b64.err; echo $?; tail -n 3 gpurun_out/b64.err
timeout 600 python bench.py --workload infer --no-cpu-baseline > gpurun_out/binf.json 2> gpurun_out/binf.err; echo $?; tail -n 3 gpurun_out/binf.err
