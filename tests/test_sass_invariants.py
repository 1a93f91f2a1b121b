"""Static checks on the compiled sm_100a code (cuobjdump -sass of libb2o.so; no GPU needed):
the hot kernels really use the Blackwell data path they claim (TMA bulk copies, tcgen05 MMA / TMEM loads), and the
ring-slot hand-back of every streaming kernel cannot overtake the shared-memory reads of the slot (the hazard found
while measuring the block apply)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "linearoperators.jl_b200", "libb2o.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    out = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, name = {}, None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = m.group(1)
            kernels[name] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m and name:
            kernels[name].append(m.group(1).strip())
    assert kernels
    return kernels


def _of(kernels, pat):
    return {k: v for k, v in kernels.items() if re.search(pat, k)}


def test_streaming_kernels_use_tma_bulk_copies(sass):
    ks = _of(sass, r"qn_compact_kernel|qn_twoloop_kernel|qn_multi_kernel")
    assert len(ks) >= 12
    for name, ins in ks.items():
        text = "\n".join(ins)
        assert "UBLKCP" in text, name                       # cp.async.bulk global -> shared
        assert "SYNCS.ARRIVE.TRANS64" in text, name         # mbarrier pipeline
        assert "SYNCS.PHASECHK.TRANS64.TRYWAIT" in text, name


def test_kron_uses_tcgen05_and_tensor_tma(sass):
    ks = _of(sass, r"kron_cluster_kernel")
    assert len(ks) == 6                                     # BM in {64, 128} x BN in {32, 64, 128}
    for name, ins in ks.items():
        text = "\n".join(ins)
        assert "UTCHMMA" in text, name                      # tcgen05.mma
        assert "UTMALDG" in text, name                      # cp.async.bulk.tensor loads
        assert "MULTICAST" in text and "UTCBAR.MULTICAST" in text, name   # TMA multicast of the shared operand, cluster-wide slot release
        assert "UTMASTG" in text, name                      # TMA-store epilogues (intermediate and result)
        assert "LDTM" in text, name                         # tcgen05.ld (TMEM -> registers)
        assert "HMMA" not in text.replace("UTCHMMA", ""), name   # no mma.sync fallback
        # the MMA / TMA issue must not sit in ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loops (operands in uniform
        # registers: one elected branch per ring stage) -- round 2 finding, profiles/r2_kron_timeline.md
        assert "R2UR.BROADCAST" not in text, name


def test_kron_pair_kernel_uses_cta_group_2(sass):
    """the batched kron kernel: tcgen05.mma.cta_group::2 issued back to back, both CTAs' tensor-map loads completing on the
    leader's barrier (.2CTA), multicast commits to both CTAs, no waterfall loop around the issue"""
    assert len(_of(sass, r"kron_pair_kernel")) == 2            # production + the instantiation with the %globaltimer stamps
    ks = _of(sass, r"kron_pair_kernelILb0")
    assert len(ks) == 1
    (name, ins), = ks.items()
    text = "\n".join(ins)
    # polling loops acquire at CTA scope (a cluster-scope acquire puts an L1 invalidate into every iteration): the only CCTL.IVALL
    # left are the cluster barriers of set-up / tear-down and the one wait whose arrivals are remote (tmem_empty)
    assert text.count("CCTL.IVALL") <= 8, text.count("CCTL.IVALL")
    assert text.count("UTCHMMA.2CTA") == 12, name              # 4 (GEMM 1) + 8 (GEMM 2: hi and lo) per ring item
    assert "UTCHMMA" not in text.replace("UTCHMMA.2CTA", ""), name
    assert text.count("UTMALDG.2D.2CTA") + text.count("UTMALDG.3D.2CTA") == 5, name
    assert text.count("UTCBAR.2CTA.MULTICAST") >= 5, name
    assert "UTMASTG" in text and "LDTM" in text, name
    assert "R2UR.BROADCAST" not in text, name
    mma = [i for i, x in enumerate(ins) if "UTCHMMA.2CTA" in x]
    runs = sorted(b - a for a, b in zip(mma, mma[1:]))
    assert runs[len(runs) // 2] <= 4, runs                     # the MMAs of an item are issued back to back


def _dest_regs(instr):
    """registers written by an LDS.{64,128}"""
    m = re.match(r"LDS(?:\.(64|128))?\s+R(\d+)", instr)
    if not m:
        return set()
    n = {None: 1, "64": 2, "128": 4}[m.group(1)]
    return {int(m.group(2)) + i for i in range(n)}


def _src_regs(instr):
    parts = instr.split(",", 1)
    if len(parts) < 2 and not instr.startswith(("STS", "ST.")):
        return set()
    body = instr if instr.startswith(("STS", "ST.", "@")) else parts[1]
    regs = set()
    for m in re.finditer(r"R(\d+)(\.64)?", body):
        regs.add(int(m.group(1)))
        regs.add(int(m.group(1)) + 1)      # 64-bit operands read a register pair
    return regs


def test_ring_slot_release_follows_the_reads_of_the_slot(sass):
    """between the last LDS.128 of a column tile and the consumer's mbarrier arrive (SYNCS.ARRIVE...A1T0) there is an
    instruction that consumes registers of that load -- so the arrive cannot issue while the load is in flight"""
    ks = _of(sass, r"qn_compact_kernel|qn_twoloop_kernel|qn_multi_kernel|qn_twoloop_multi_kernel")
    checked, bad = 0, []
    for name, ins in ks.items():
        for i, instr in enumerate(ins):
            if "SYNCS.ARRIVE.TRANS64.A1T0" not in instr:
                continue
            j = i - 1
            while j >= 0 and not re.match(r"(@!?U?P\d+\s+)?LDS\.128", ins[j]):
                j -= 1
            if j < 0 or i - j > 400:
                continue                                   # arrive not related to a tile read (e.g. prologue)
            dest = _dest_regs(re.sub(r"^@!?U?P\d+\s+", "", ins[j]))
            used = any(dest & _src_regs(re.sub(r"^@!?U?P\d+\s+", "", x)) for x in ins[j + 1:i])
            if not used:
                bad.append("%s: arrive at +%d right behind an unconsumed %s" % (name, i, ins[j]))
            checked += 1
    assert not bad, "\n".join(bad)
    assert checked >= 30


def test_sparse_kernels_are_atomics_free_and_the_tile_kernel_uses_tma(sass):
    """the sparse-matrix leaf claims fixed-order, atomics-free sums (bit-reproducible) in all three kernels, and that the tile
    kernel stages idx / val / offsets with bulk copies behind an mbarrier ring; the row kernels issue the
    loads of a whole 4-wide trip (idx, val and the four gathers: >= 8 read-only global loads) before the first multiply-add"""
    ks = _of(sass, r"spmv_rows_kernel|spmv_rows_pipe_kernel|spmv_tiles_kernel")
    assert len(ks) == 3 * 6 * 2                              # 3 kernels x 6 lane-group widths x {Float64, Float32}
    for name, ins in ks.items():
        text = "\n".join(ins)
        assert not re.search(r"\b(ATOM|ATOMG|ATOMS|RED)\b", text), name
        assert "DFMA" in text, name                          # sums are taken in double for both element types
        if "spmv_tiles_kernel" in name:
            assert text.count("UBLKCP") >= 3, name           # idx, val, offsets
            assert "SYNCS.ARRIVE.TRANS64" in text and "SYNCS.PHASECHK.TRANS64.TRYWAIT" in text, name
        else:
            first = next(i for i, x in enumerate(ins) if "DFMA" in x)
            loads = sum(1 for x in ins[max(0, first - 60):first] if re.search(r"LDG\.E(\.64)?\.CONSTANT", x))
            assert loads >= 8, (name, loads)               # 4 (idx, val) pairs' worth of loads + the 4 gathers, before any multiply
