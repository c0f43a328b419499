#!/bin/bash
# Round-end measurement on ONE B200: GPU tests, smoke, bench, ncu launch list and full captures of the hot kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ "$1" != "swe" ]; then
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err && echo BENCH_OK
python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_lb_fast|k_gather_chunks' -s 8 -c 2 -o gpurun_out/prof_lb -f \
    python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_lb.log 2>&1
GG_LB_CLASSES=1 python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/plain_cls.log 2>&1 && \
GG_LB_CLASSES=1 ncu --set full --clock-control none --import-source on -k regex:'k_lb_fast|k_gather_chunks|k_lb_residual' -s 12 -c 3 -o gpurun_out/prof_lb_cls -f \
    python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_lb_cls.log 2>&1
fi
if [ "$1" != "swe" ]; then ls -la gpurun_out | tail -20; exit 0; fi
rm -f gpurun_out/prof_lb.ncu-rep gpurun_out/prof_lb_cls.ncu-rep   # the merged directory is capped at 64 MiB
python tools/swe_profile.py --n 256 --order 2 --steps 1 > gpurun_out/plain_swe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_pcg_ebe|k_swe_diag|k_swe_stage|k_ebe_scalar_mass' -s 30 -c 8 -o gpurun_out/prof_swe -f \
    python tools/swe_profile.py --n 256 --order 2 --steps 1 > gpurun_out/ncu_swe.log 2>&1
ls -la gpurun_out | tail -20
