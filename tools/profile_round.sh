#!/bin/bash
# One GPU-box pass for the round's evidence (run under gpurun from the repo root): tests, bench lines per named config,
# ncu launch list of the bench command, per-kernel figures of the block coder, full captures of the hot kernels.
# Outputs under gpurun_out/<tag>_*; tools/make_profiles.py turns them into profiles/.
tag=${1:-r02f}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${tag}_tests.log
python bench.py --config 2 > gpurun_out/${tag}_bench_config2.json 2> gpurun_out/${tag}_bench_config2.err; echo "bench c2 rc=$?"
python bench.py --config 1 --no-cli > gpurun_out/${tag}_bench_config1.json 2> gpurun_out/${tag}_bench_config1.err; echo "bench c1 rc=$?"
python bench.py --config 3 --no-cli > gpurun_out/${tag}_bench_config3.json 2> gpurun_out/${tag}_bench_config3.err; echo "bench c3 rc=$?"
python bench.py --config 5 --no-cli > gpurun_out/${tag}_bench_config5.json 2> gpurun_out/${tag}_bench_config5.err; echo "bench c5 rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "ref rc=$?"
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --config 2 --steps 1 --warmup 1 --no-cpu > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "launch list rc=$?"
# block coder: instructions and issue-slot utilisation per launch, one resident encode + decode x 4
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"k2_model|k2_code|k2_coder|k2_block" -c 48 --csv --log-file gpurun_out/${tag}_k2_launches.csv \
    python tools/sweep_blocks.py 1 auto 1 > gpurun_out/${tag}_ncu_k2l.log 2>&1; echo "k2 list rc=$?"
# full captures (one launch each, the last generation's): model kernel, interval kernel, decoder, K1 (tail launch), K3
ncu --set full --import-source on --clock-control none -k regex:"k2_model|k2_code" -s 8 -c 2 -o gpurun_out/${tag}_k2enc -f \
    python tools/sweep_blocks.py 1 auto 1 > gpurun_out/${tag}_ncu_k2enc.log 2>&1; echo "k2enc rc=$?"
ncu --set full --import-source on --clock-control none -k regex:"k2_coder" -s 4 -c 1 -o gpurun_out/${tag}_k2dec -f \
    python tools/sweep_blocks.py 1 auto 1 > gpurun_out/${tag}_ncu_k2dec.log 2>&1; echo "k2dec rc=$?"
ncu --set full --import-source on --clock-control none -k regex:"k1_extract|k3_reconstruct" -s 1 -c 2 -o gpurun_out/${tag}_k1k3 -f \
    python tools/sweep_blocks.py 1 auto 1 > gpurun_out/${tag}_ncu_k1k3.log 2>&1; echo "k1k3 rc=$?"
ls -la gpurun_out | grep ${tag} | awk '{print $5, $9}'
