set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r01_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gi_small -s 1 -c 1 -o gpurun_out/r01_gi_small_c2 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/b2.log 2>&1
ncu --set full --clock-control none -k regex:k1_psi_fill -c 1 -o gpurun_out/r01_psi_fill -f python -c "
import sys; sys.path.insert(0,'.')
from copra_b200 import capi, workloads as wl
e=capi.Engine(0); bp=wl.c5(batch=64); hb=capi.HostBatch(bp); e.lmpc_build(hb); p=e.download(hb,'Psi'); print(p.shape)
" > gpurun_out/b3.log 2>&1
tail -c 600 gpurun_out/bench_c2.json
