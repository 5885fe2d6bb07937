"""Static checks on the CUDA sources that no GPU run can give cheaply.

Programmatic dependent launch (csrc/launch.h): a kernel launched through launch_k() may
start while its predecessor is still running; it is only correct if EVERY thread executes
griddepcontrol.wait (pdl_enter() / pdl_wait()) before its first memory access.  A kernel
launched that way without the wait would race silently -- so the rule is checked here."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "multigrid_parallel_b200", "csrc")


def _sources():
    return {f: open(os.path.join(CSRC, f)).read() for f in sorted(os.listdir(CSRC))
            if f.endswith((".cu", ".cuh", ".h"))}


def _kernel_bodies(text):
    """name -> body text of every __global__ function in a source file"""
    out = {}
    for m in re.finditer(r"__global__[^{;]*?\b(k_\w+)\s*\([^{;]*?\)\s*\{", text, re.S):
        depth, i = 1, m.end()
        while depth and i < len(text):
            depth += {"{": 1, "}": -1}.get(text[i], 0)
            i += 1
        out.setdefault(m.group(1), []).append(text[m.end():i])
    return out


def test_every_pdl_launched_kernel_waits_before_touching_memory():
    src = _sources()
    launched = set()
    for text in src.values():
        launched |= set(re.findall(r"\blaunch_k\(\s*(k_\w+)", text))
    assert len(launched) >= 20, launched  # the cycle's kernels all go through launch_k
    bodies = {}
    for text in src.values():
        for name, bs in _kernel_bodies(text).items():
            bodies.setdefault(name, []).extend(bs)
    for name in sorted(launched):
        assert name in bodies, f"{name} is launched through launch_k but not defined"
        for body in bodies[name]:
            m = re.search(r"\bpdl_(enter|wait)\(\)", body)
            assert m, f"{name}: launched with the PDL attribute but never waits"
            before = body[:m.start()]
            # nothing that reads or writes memory in front of the wait: no dereference of a
            # kernel argument, no load/store helper, no atomics
            assert not re.search(r"__ld|__st|atomic|tma_load|tma_prefetch|ld2\(|st2\(|P\.\w+\[", before), \
                f"{name}: memory access before the PDL wait:\n{before[-300:]}"
            # ... and every early return in front of it waits first
            for r in re.finditer(r"\breturn\b", before):
                window = before[max(0, r.start() - 120):r.start()]
                assert "pdl_wait()" in window, f"{name}: returns before the PDL wait"


def test_kernels_without_the_wait_use_plain_launches():
    """the converse: <<<...>>> launches (vtk.cu, gslex.cu, set-up kernels) never carry the
    attribute, so kernels without a wait are safe by construction; just make sure nobody
    wraps cudaLaunchKernelEx elsewhere"""
    for f, text in _sources().items():
        if f != "launch.h":
            assert "cudaLaunchKernelEx" not in text, f
            assert "ProgrammaticStreamSerialization" not in text, f


def test_product_does_not_reference_the_oracle():
    for base in ("multigrid_parallel_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "_build" in dirpath or "__pycache__" in dirpath:
                continue
            for f in files:
                if f.endswith((".so", ".pyc", ".o")):
                    continue
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"oracle_lib|liborc|mg_oracle|oracle/", text), os.path.join(dirpath, f)
