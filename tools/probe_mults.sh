# e2e encode / decode wall clock for a few per-chunk block-size multipliers (run under gpurun): bash tools/probe_mults.sh
for m in "1.3,1.2,1.0,0.6,0.45" "1.3,1.2,1.0,0.7,0.45" "1.4,1.25,1.0,0.7,0.45" "1.35,1.25,1.0,0.65,0.4"; do
    echo "MULTS $m"; CBCG_PIPE_MULTS=$m python tools/e2e_probe.py 1 default 2>/dev/null
done
