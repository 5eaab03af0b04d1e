# The library's host code (tile planner, partition, Matrix Market ingest, generators, solver set-up and
# loops, peer-memory set-up and teardown) under AddressSanitizer + UBSan, on the device stand-in of
# tests/hostsim (no GPU involved; the CUDA kernels are outside its reach).
#   bash tools/asan_hostsim.sh
cd "$(dirname "$0")/.."
set -e
D=tests/hostsim/build_asan
mkdir -p $D
FL="-O1 -g -std=gnu11 -fPIC -fopenmp -fsanitize=address,undefined -fno-omit-frame-pointer -Iinclude -Iacg_b200/csrc -Itests/hostsim -I/usr/local/cuda/include"
for f in error vector symcsrmatrix stencil rmat mtxfile metis_rows comm halo p2p compress slices mergeplan plan cgcuda expand_host ext; do /usr/bin/gcc $FL -c acg_b200/csrc/$f.c -o $D/$f.o; done
for f in cuda_mock nccl_mock kernels_sim; do /usr/bin/gcc $FL -c tests/hostsim/$f.c -o $D/$f.o; done
/usr/bin/gcc -shared -Wl,-Bsymbolic -fsanitize=address,undefined -o $D/libacgb200_hostsim.so $D/*.o -lgomp -lpthread -lrt -lm
export LD_PRELOAD="$(/usr/bin/gcc -print-file-name=libasan.so) $(/usr/bin/gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0 OMP_NUM_THREADS=2 ACGB200_P2P_TIMEOUT_MS=60000
set +e
echo "== host structures, ingest, generators"
ACGB200_TEST_LIB=$PWD/$D/libacgb200_hostsim.so python -m pytest tests/test_host_structs.py tests/test_mtxfile.py -q -x \
    -k "not metis and not thread_count and not reference and not refuses" -p no:cacheprovider 2>&1 | tail -3
echo "== solver, 3 ranks, every loop back-end"
ACGB200_TEST_HOSTSIM=$PWD/$D/libacgb200_hostsim.so python -m torch.distributed.run --nnodes=1 --nproc-per-node=3 \
    --master-addr 127.0.0.1 --master-port 31114 tests/_dist_worker.py --mode gpu --matrix 27pt --size 16 --partition block \
    --backends p2p-fused,p2p-unfused,tiles-only,nccl,nccl-graph 2>&1 \
    | grep -c " OK$\|FAIL\|runtime error\|AddressSanitizer" 
echo "== every runtime call of set-up and solves failing in turn (159 at the end of round 1)"
for k in $(seq 1 260); do ACGB200_TEST_LIB=$PWD/$D/libacgb200_hostsim.so HOSTSIM_FAIL_CALL_AT=$k python tests/hostsim/run_fault.py 2>&1 | grep -i "Sanitizer\|runtime error\|^ok\|^error"; done | sort | uniq -c
rm -f /dev/shm/acgb200nccl_* /dev/shm/acgb200sim_*
