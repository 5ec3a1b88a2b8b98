"""The C program tests/c/replay_julia_calls.c replays the exact ccall sequence of every method override in
semiseparablematrices.jl_b200/julia/SemiseparableB200.jl (same entry points, argument order, strides, nb = 1) and checks the results against dense
arithmetic.  CPU tier: it compiles and links with gcc against include/ssm_b200.h alone (the boundary is a C ABI).  GPU tier: it runs."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIBDIR = ROOT / "semiseparablematrices.jl_b200" / "lib"
SRC = ROOT / "tests" / "c" / "replay_julia_calls.c"
EXE = ROOT / "tests" / "c" / "replay_julia_calls"


def build():
    subprocess.run(["gcc", "-O1", "-std=c11", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(SRC), "-o", str(EXE), "-L", str(LIBDIR),
                    "-lssmb200", "-lm", f"-Wl,-rpath,{LIBDIR}"], check=True)


def test_replay_program_builds_against_the_c_header_only():
    build()
    assert EXE.exists()


def test_every_ccall_of_the_julia_glue_is_replayed():
    """each `ccall((:ssm_xxx, libssm), ...)` symbol of the .jl appears in the C replay (so a signature change breaks one of the two visibly)"""
    jl = (ROOT / "semiseparablematrices.jl_b200" / "julia" / "SemiseparableB200.jl").read_text()
    called = set(re.findall(r"ccall\(\(:(ssm_[a-z0-9_]+),", jl))
    replayed = set(re.findall(r"\b(ssm_[a-z0-9_]+)\s*\(", SRC.read_text()))
    bookkeeping = {"ssm_pool_trim", "ssm_bps_get", "ssm_bpsqr_rq_mul"}       # pool control / the rq step have their own GPU tests (test_gpu_bps.py, test_gpu_edge.py)
    assert called - replayed - bookkeeping == set(), called - replayed - bookkeeping


@pytest.mark.gpu
def test_replay_runs_on_the_gpu():
    build()
    p = subprocess.run([str(EXE)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "replay OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
