python -c "import torch"
for cfg in "1212416 64" "2097152 64" "3276800 10" "3276800 64" "303104 64"; do set -- $cfg; echo "== M=$1 S=$2"; BF_S=$2 PNB_FUSED_PROF=1 timeout 120 python tools/bench_fused.py $1 2>&1 | grep -E "cycles/CTA|kernel" | tail -2 | cut -c1-230; done
