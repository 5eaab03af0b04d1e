"""compat/mpi: the MPI stand-in that lets the unmodified reference driver run as several processes on
an image without MPI (CPU only).  tests/mpishim/shim_selftest.c exercises point-to-point (head-to-head
64 MB exchanges, tags out of order, non-blocking), every collective the reference's host layer calls,
MPI_IN_PLACE, communicator contexts and contiguous datatypes; it aborts on the first mismatch."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MPI = os.path.join(ROOT, "compat", "mpi")


@pytest.fixture(scope="module")
def selftest(tmp_path_factory):
    p = subprocess.run(["make", "-C", MPI], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    exe = str(tmp_path_factory.mktemp("mpishim") / "shim_selftest")
    cc = subprocess.run(["/usr/bin/gcc", "-O2", "-std=gnu11", "-Wall", "-I" + MPI, os.path.join(ROOT, "tests", "mpishim", "shim_selftest.c"),
                         "-o", exe, "-L" + MPI, "-lacgb200mpishim", "-Wl,-rpath," + MPI, "-lm"], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr[-2000:]
    return exe


@pytest.mark.parametrize("nranks", [1, 2, 5, 8])
def test_selftest_under_the_launcher(nranks, selftest):
    p = subprocess.run([os.path.join(MPI, "acgb200-mpirun"), "-n", str(nranks), selftest], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == f"ok {nranks} ranks", p.stdout[-500:] + p.stderr[-2000:]


def test_single_process_without_launcher(selftest):
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "ACGB200_MPI_RANK", "ACGB200_MPI_SIZE")}
    p = subprocess.run([selftest], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == "ok 1 ranks"


def test_torchrun_style_environment(selftest):
    """Any launcher that exports RANK / WORLD_SIZE / MASTER_PORT works (e.g. torchrun --no-python)."""
    procs = [subprocess.Popen([selftest], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE="3", MASTER_PORT=str(40000 + os.getpid() % 20000)))
             for r in range(3)]
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert outs[0][0].strip() == "ok 3 ranks"


def test_a_failing_rank_ends_the_job(tmp_path):
    src = tmp_path / "abort.c"
    src.write_text("""#include <mpi.h>
#include <unistd.h>
int main(int argc, char **argv)
{
    int rank;
    MPI_Init(&argc, &argv);
    MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    if (rank == 1) MPI_Abort(MPI_COMM_WORLD, 7);
    MPI_Barrier(MPI_COMM_WORLD);
    sleep(60);
    MPI_Finalize();
    return 0;
}
""")
    exe = str(tmp_path / "abort")
    subprocess.run(["make", "-C", MPI], check=True, capture_output=True)
    subprocess.run(["/usr/bin/gcc", "-I" + MPI, str(src), "-o", exe, "-L" + MPI, "-lacgb200mpishim", "-Wl,-rpath," + MPI], check=True)
    p = subprocess.run([os.path.join(MPI, "acgb200-mpirun"), "-n", "3", exe], capture_output=True, text=True, timeout=60)
    assert p.returncode != 0
