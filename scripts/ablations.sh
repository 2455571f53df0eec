# Ablations of the step-level mechanisms at N=1 (graph replay): run on a B200 via gpurun.
run() { timeout 200 python bench.py --no-cpu-baseline --steps 60 --warmup 10 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['ms_per_step'],3))"; }
run default
MMNAS_PDL=0 run no_pdl
MMNAS_GEMM_PAIR=0 run no_cta_pair
MMNAS_OVERLAP_WGRAD=0 run no_side_stream
