# Round 2: the throughput-bound regime (16384 arenas on one GPU, steady state): do the instruction-saving island mappings that lose at
# 4096 arenas (chain-bound) win here?
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { python bench.py --arenas 16384 --steps 12 --warmup 100 --no-cpu-baseline --no-extras --no-parity --no-steady --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in r['world_step_stage_ms_per_tick'].items()}, {k: round(r['kernel_ms_total'][k]/max(1,r['kernel_launch_counts'][k]),2) for k in ('sense','think','isl_big','isl_warp','isl_batch','isl_lanex')})"; }
for e in "PS_X=0" "PS_LANEX_MAX_CONTACTS=27" "PS_LANEX_MAX_CONTACTS=27 PS_BATCHX_BLOCKS_PER_SM=2" "PS_ISLAND_G=8" "PS_ISLAND_G=8 PS_POS_BLOCKS_PER_SM=8" "PS_ISLAND_G=4 PS_POS_BLOCKS_PER_SM=4" "PS_ISLAND_G=8 PS_ISLAND_SPLIT=1 PS_POS_BLOCKS_PER_SM=8" "PS_BATCH_BLOCKS_PER_SM=3" "PS_WORKER_BLOCKS_PER_SM=6"; do
  echo "variant: $e"; env $e bash -c "$(declare -f run); run"
done 2>&1 | tee gpurun_out/r2_ab_bigbatch.log
