#!/bin/bash
# Build the UNMODIFIED reference driver cuda/acg-cuda.c against libacgb200_mpi.so.
#
# Everything compiled here comes from /root/reference where it lies (nothing is
# copied into the repository); outputs go to oracle/_ref/driver/ (git-ignored,
# shipped to the GPU box by gpurun).  The reference's CUDA solver sources
# (acg/cgcuda.c, acg/cg-kernels-cuda.cu, acg/halo.cu, acg/comm.c and the NVSHMEM
# wrappers) are NOT compiled: their symbols come from the library.  mpi.h is the
# stand-in compat/mpi/mpi.h (libacgb200mpishim.so): one rank when started directly, N ranks on one
# node under compat/mpi/acgb200-mpirun -n N.
set -e
REF=${REF:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/oracle/_ref/driver
CUDA=${CUDA:-/usr/local/cuda}
mkdir -p "$OUT"
make -s -C "$ROOT/acg_b200/csrc" mpi
DEFS="-D_GNU_SOURCE -DHAVE_CLOCK_GETTIME -DACG_HAVE_OPENMP -DACG_HAVE_MPI -DACG_HAVE_NCCL -DACG_HAVE_CUBLAS -DACG_HAVE_CUSPARSE"
INC="-I$REF -I$ROOT/compat/mpi -I$CUDA/include"
# METIS: the static archive inside the CUDA toolkit + the interface header compat/metis/metis.h, so the driver
# partitions rows as the reference does (acg/metis.c:225-346) when no --partition file is given
METIS_LIB="$CUDA/targets/x86_64-linux/lib/libmetis_static.a"
if [ -f "$METIS_LIB" ]; then DEFS="$DEFS -DACG_HAVE_METIS"; INC="$INC -I$ROOT/compat/metis"; else METIS_LIB=""; fi
/usr/bin/gcc -O2 -g -fopenmp $DEFS -DACG_HAVE_CUDA $INC -c "$REF/cuda/acg-cuda.c" -o "$OUT/acg-cuda.o"
# host layer of the reference, unchanged; halo.c without ACG_HAVE_CUDA = pattern code only;
# cgpetsc.c without PETSc = the reference's own 'not supported' stubs
for f in vector symcsrmatrix graph halo error fmtspec mtxfile metis prefixsum sort cgpetsc; do
  /usr/bin/gcc -O2 -g -fopenmp $DEFS $INC -c "$REF/acg/$f.c" -o "$OUT/$f.o"
done
/usr/bin/gcc -fopenmp -o "$OUT/acg-cuda" "$OUT"/*.o \
  -L"$ROOT/acg_b200" -lacgb200_mpi -Wl,-rpath,'$ORIGIN/../../../acg_b200' \
  -L"$ROOT/compat/mpi" -lacgb200mpishim -Wl,-rpath,'$ORIGIN/../../../compat/mpi' \
  $METIS_LIB -L"$CUDA/lib64" -lcublas -lcusparse -lcudart -lnccl -lm
echo "built $OUT/acg-cuda"

# The stock reference GPU solver (its own cgcuda.c: cusparseSpMV + cublasDdot + its axpy kernels) with the same
# single-process MPI stand-in, for sm_100a: the "kernel to beat" arm (bench.py --with-reference-gpu).
REFOUT=$ROOT/oracle/_ref/driver_ref
mkdir -p "$REFOUT"
/usr/bin/gcc -O2 -fopenmp $DEFS -DACG_HAVE_CUDA $INC -c "$REF/cuda/acg-cuda.c" -o "$REFOUT/acg-cuda.o"
for f in vector symcsrmatrix graph halo comm error fmtspec mtxfile metis prefixsum sort cgpetsc cgcuda cg; do
  /usr/bin/gcc -O2 -fopenmp $DEFS -DACG_HAVE_CUDA $INC -c "$REF/acg/$f.c" -o "$REFOUT/$f.o"
done
for f in cg-kernels-cuda halo comm-nvshmem nvshmem; do
  "$CUDA/bin/nvcc" -O3 -gencode arch=compute_100a,code=sm_100a -rdc=true -Xcompiler -fopenmp $DEFS -DACG_HAVE_CUDA $INC \
      -c "$REF/acg/$f.cu" -o "$REFOUT/$f.cu.o"
done
"$CUDA/bin/nvcc" -gencode arch=compute_100a,code=sm_100a -Xcompiler -fopenmp -o "$REFOUT/acg-cuda-ref" "$REFOUT"/*.o \
  -L"$ROOT/compat/mpi" -lacgb200mpishim -Xlinker -rpath -Xlinker '$ORIGIN/../../../compat/mpi' \
  $METIS_LIB -lcublas -lcusparse -lnccl -lm -lcudadevrt
echo "built $REFOUT/acg-cuda-ref"
