"""The shm MPI stand-in used ONLY to build/time the unmodified reference (baseline/shim): ring Put into a window,
in-place SUM all-reduce (chunked), synchronous Issend/Recv/Wait -- exactly the calls SURVEY.md section 2.5 lists."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "baseline", "shim")

PROG = r"""
#include "mpi.h"
#include <stdio.h>
#include <stdlib.h>
int main(int argc, char **argv) {
    int rank, size; MPI_Win win; float *mem;
    MPI_Init(&argc, &argv); MPI_Comm_size(MPI_COMM_WORLD, &size); MPI_Comm_rank(MPI_COMM_WORLD, &rank);
    int n = 1000, left = (rank + size - 1) % size, right = (rank + 1) % size;
    MPI_Alloc_mem(2 * n * sizeof(float), MPI_INFO_NULL, &mem);
    MPI_Win_create(mem, 2 * n * sizeof(float), sizeof(float), MPI_INFO_NULL, MPI_COMM_WORLD, &win);
    for (int i = 0; i < 2 * n; i++) mem[i] = 0;
    float *buf = malloc(n * sizeof(float));
    for (int i = 0; i < n; i++) buf[i] = rank * 1000 + i;
    MPI_Barrier(MPI_COMM_WORLD);
    MPI_Win_lock(MPI_LOCK_SHARED, left, 0, win); MPI_Put(buf, n, MPI_FLOAT, left, n, n, MPI_FLOAT, win); MPI_Win_unlock(left, win);
    MPI_Win_lock(MPI_LOCK_SHARED, right, 0, win); MPI_Put(buf, n, MPI_FLOAT, right, 0, n, MPI_FLOAT, win); MPI_Win_unlock(right, win);
    MPI_Barrier(MPI_COMM_WORLD);
    int bad = 0;
    for (int i = 0; i < n; i++) { if (mem[i] != left * 1000 + i) bad++; if (mem[n + i] != right * 1000 + i) bad++; }
    /* chunked all-reduce: 3M floats > 4 MiB staging */
    int big = 3000000; float *g = malloc(big * sizeof(float));
    for (int i = 0; i < big; i++) g[i] = (float)(rank + 1);
    MPI_Allreduce(MPI_IN_PLACE, g, big, MPI_FLOAT, MPI_SUM, MPI_COMM_WORLD);
    float want = size * (size + 1) / 2.0f;
    for (int i = 0; i < big; i += 997) if (g[i] != want) bad++;
    int ev = 2; MPI_Allreduce(MPI_IN_PLACE, &ev, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD); if (ev != 2 * size) bad++;
    /* decent-style two-sided exchange */
    MPI_Request r1, r2; float *lb = malloc(n * sizeof(float)), *rb = malloc(n * sizeof(float));
    MPI_Issend(buf, n, MPI_FLOAT, left, 10, MPI_COMM_WORLD, &r1);
    MPI_Issend(buf, n, MPI_FLOAT, right, 2, MPI_COMM_WORLD, &r2);
    MPI_Recv(lb, n, MPI_FLOAT, left, 2, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    MPI_Recv(rb, n, MPI_FLOAT, right, 10, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    MPI_Wait(&r1, MPI_STATUS_IGNORE); MPI_Wait(&r2, MPI_STATUS_IGNORE);
    for (int i = 0; i < n; i++) { if (lb[i] != left * 1000 + i) bad++; if (rb[i] != right * 1000 + i) bad++; }
    double t = MPI_Wtime(); if (t <= 0) bad++;
    printf("rank %d bad %d\n", rank, bad);
    MPI_Finalize();
    return bad != 0;
}
"""


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
@pytest.mark.parametrize("world", [1, 2, 3])
def test_shim_put_allreduce_sendrecv(tmp_path, world):
    src = tmp_path / "t.c"
    src.write_text(PROG)
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-O1", "-I", SHIM, str(src), os.path.join(SHIM, "egmpi.c"), "-o", str(exe), "-lpthread", "-lrt"])
    p = subprocess.run([sys.executable, os.path.join(SHIM, "mpirun"), "-np", str(world), str(exe)],
                       capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("bad 0") == world
