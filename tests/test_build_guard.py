"""Guards for the register-bound E-step kernels: the sweep loops sit at the
255-register limit, and an edit next to them can tip ptxas into spilling tile data or into an
allocation full of register moves.  Runs on the objects the build leaves under csrc/build (no GPU)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mrlda_b200", "csrc")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("nvcc") is None,
                                reason="CUDA toolkit not installed")


@pytest.fixture(scope="module")
def objects():
    subprocess.check_call(["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL, timeout=1500)
    return os.path.join(CSRC, "build")


def res_usage(obj):
    out = subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True, check=True).stdout
    res = {}
    for fn, reg, stack in re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        res[fn] = (int(reg), int(stack))
    return res


def sass(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    fns, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        if "Function :" in line:
            cur = line.split(":", 1)[1].strip()
            fns[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            fns[cur].append((int(m.group(1), 16), m.group(2)))
    return fns


def loops(ins):
    """(start, end) instruction index pairs of the backward branches."""
    idx = {a: i for i, (a, _) in enumerate(ins)}
    out = []
    for i, (a, t) in enumerate(ins):
        if re.search(r"\bBRA\b", t):
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in idx:
                out.append((idx[int(m.group(1), 16)], i))
    return out


def is_move(t):
    return "IMAD.MOV" in t or re.match(r"(@!?U?P\d+\s+)?MOV\b", t) is not None


@pytest.mark.parametrize("obj", ["estep3_inst_L8o.o", "estep3_inst_L8.o", "estep3_inst_L8n17.o", "estep3_inst_L8n20.o",
                                 "estep3_inst_L8n22.o"])
def test_warp_group_kernel_sweeps_do_not_spill(objects, obj):
    path = os.path.join(objects, obj)
    usage = {f: v for f, v in res_usage(path).items() if "estep3_kernel" in f}
    assert len(usage) == 4  # G = 1, 2, 4, 8
    for fn, (reg, stack) in usage.items():
        assert reg <= 255 and stack <= 64, (fn, reg, stack)  # a few cold per-document values at most
    checked = 0
    for fn, ins in sass(path).items():
        if "estep3_kernel" not in fn:
            continue
        # the specialised sweep loops: straight-line bodies with the group barriers inside
        sweeps = [(s, e) for s, e in loops(ins) if 300 < e - s < 2500 and
                  any("BAR.SYNC" in t for _, t in ins[s:e + 1])]
        for s, e in sweeps:
            body = [t for _, t in ins[s:e + 1]]
            assert not any(re.search(r"\b(LDL|STL)\b", t) for t in body), (fn, hex(ins[s][0]))
            moves = sum(1 for t in body if is_move(t))
            # ~80 moves of fixed work (constants of exp(digamma), the S exchange) plus a few per batch
            # (the narrower-row builds for 105..176 topics are not hand-tuned: a little more slack)
            slack = 60 if obj in ("estep3_inst_L8o.o", "estep3_inst_L8.o") else 80
            assert moves <= slack + 0.10 * len(body), (fn, hex(ins[s][0]), moves, len(body))
            checked += 1
    assert checked >= 20


@pytest.mark.parametrize("obj,spill_max", [("estep3c_inst_C5o.o", 2), ("estep3c_inst_C2.o", 8)])
def test_cluster_warp_group_kernel_sweeps_do_not_spill(objects, obj, spill_max):
    """The same loops with the cross-CTA norm exchange in them (estep3c_kernel.cuh): at most a stray
    scalar in local memory in cfg 5's K = 1000 build (C5o), a few in the 26-column builds."""
    path = os.path.join(objects, obj)
    usage = {f: v for f, v in res_usage(path).items() if "estep3c_kernel" in f}
    assert len(usage) == 4  # G = 1, 2, 4, 8
    for fn, (reg, stack) in usage.items():
        assert reg <= 255 and stack <= 64, (fn, reg, stack)
    checked = 0
    for fn, ins in sass(path).items():
        if "estep3c_kernel" not in fn:
            continue
        sweeps = [(s, e) for s, e in loops(ins) if 300 < e - s < 3200 and
                  any("BAR.SYNC" in t for _, t in ins[s:e + 1])]
        for s, e in sweeps:
            body = [t for _, t in ins[s:e + 1]]
            assert sum(1 for t in body if re.search(r"\b(LDL|STL)\b", t)) <= spill_max, (fn, hex(ins[s][0]))
            assert any("SYNCS" in t or "ST.ASYNC" in t.upper() or "STAS" in t for t in body), (fn, hex(ins[s][0]))
            checked += 1
    assert checked >= 20


def test_register_stationary_kernel_allocation(objects):
    """estep2_kernel<8,26,2,8,1> (K = 200 when MRLDA_ESTEP_KERNEL=2, and the template of the cluster
    variants): 250-255 registers, no spills, the fast register allocation (the
    slow one showed ~935 IMAD.MOV)."""
    path = os.path.join(objects, "estep2_inst_W8.o")
    key = "estep2_kernelILi8ELi26ELi2ELi8ELi1E"
    usage = [v for f, v in res_usage(path).items() if key in f]
    assert usage and all(reg <= 255 and stack == 0 for reg, stack in usage), usage
    ins = [i for f, i in sass(path).items() if key in f][0]
    text = [t for _, t in ins]
    assert not any(re.search(r"\b(LDL|STL)\b", t) for t in text)
    assert sum(1 for t in text if "IMAD.MOV" in t) <= 650
