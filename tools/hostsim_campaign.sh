# Host-logic campaign on the device stand-in (tests/hostsim): matrices x partitions x rank counts x
# every loop back-end, against the single-rank oracle.  CPU only; a few minutes.
#   bash tools/hostsim_campaign.sh
cd "$(dirname "$0")/.."
make -C tests/hostsim >/dev/null || exit 1
export ACGB200_TEST_HOSTSIM=$PWD/tests/hostsim/libacgb200_hostsim.so OMP_NUM_THREADS=2 ACGB200_P2P_TIMEOUT_MS=20000
fail=0
for cfg in "2 7pt 7 slab" "4 27pt 9 block" "3 rmat 2000 random" "4 7pt 10 random" "5 27pt 10 slab" "8 27pt 12 block"; do
  set -- $cfg
  extra=""; [ "$2" = "rmat" ] && extra="--maxits 10 --rtol 0"
  out=$(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port $((30000 + RANDOM % 2000)) \
        tests/_dist_worker.py --mode gpu --matrix $2 --size $3 --partition $4 \
        --backends p2p-fused,p2p-unfused,tiles-only,nccl,nccl-graph $extra 2>&1)
  ok=$(echo "$out" | grep -c " OK$"); bad=$(echo "$out" | grep -c "FAIL\|Traceback")
  echo "$cfg: ok=$ok bad=$bad"
  [ "$bad" != "0" ] && { echo "$out" | grep "FAIL\|Error" | head -5; fail=1; }
done
rm -f /dev/shm/acgb200nccl_* /dev/shm/acgb200sim_*
exit $fail
