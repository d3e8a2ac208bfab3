"""Sweep planner: gate list -> sequence of *sweeps* over the statevector, each a sequence of register
*passes* -- the static program the CUDA sweep kernel interprets (csrc/qb_kernels.cuh).

Terminology
  tile     2^k amplitudes (k = TILE_BITS) that one CTA holds on chip; the tile's bit positions are k
           qubits ``tile_qubits`` (ascending; always containing the LOW_BITS lowest qubits so that every
           global access is a run of 2^LOW_BITS contiguous amplitudes = 256 B for complex128).
  sweep    one read + one write of the whole state (2 * 16 B * 2^n for complex128): the unit of HBM
           traffic.  All gates of a sweep have their *target* among the tile qubits; a control may be any
           qubit (a control outside the tile is a per-tile predicate).
  pass     within a sweep, each thread holds 2^r amplitudes (r = REG_BITS) in registers, spanning r tile
           bits ``reg_bits``; gates targeting those bits are applied in registers.  Passes exchange data
           through shared memory; the first pass reads global memory directly and the last pass writes
           it directly, so both must keep the LOW_BITS lowest tile bits as thread (lane) bits to stay
           coalesced.

Ordering rule: two gates may be reordered iff on every shared qubit both act diagonally (controls and
DIAG targets).  The planner is a greedy list scheduler under that rule; it never changes the product of
the circuit's unitaries; reordering commuting gates changes results only at the level of fp rounding.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from .gate_list import DENSE, DIAG, KernelOp, dfma_per_amplitude

TILE_BITS = 11  # default; 12 is also compiled (include/queasars_b200.h QB_MAX_TILE_BITS)
REG_BITS = 4
LOW_BITS = 4
MAX_SWEEP_OPS = 96  # csrc/qb_kernels.cuh kMaxSweepOps
MAX_SWEEP_PASSES = 16  # kMaxSweepPasses

PASS_FLAG_WARP_LOCAL = 1  # qb_pass.flags bit 0
# the first / last pass of a sweep may hold low (lane) bits in registers -- no empty edge passes, but the HBM accesses of those
# passes are only partly coalesced (measured +1.8 % with the restart planner; 0 = fully coalesced edge passes)
ALLOW_LOW_EDGE_PASSES = os.environ.get("QB_ALLOW_LOW_EDGE", "1") != "0"
PLAN_TRIALS = int(os.environ.get("QB_PLAN_TRIALS", "48"))  # randomised restarts of the sweep (tile) choice; 0 = greedy only
PLAN_FULL_BUILDS = int(os.environ.get("QB_PLAN_FULL_BUILDS", "3"))  # how many of the best draws are planned in full (passes) before the winner is chosen
PLAN_VISIT_BUDGET = 400_000  # op visits the restarts may spend per circuit (48 trials up to ~8 000 op-sweeps)
PLAN_ACCEPT = (0.9, 0.8, 0.7)  # probability of accepting a new tile qubit in a randomised trial (cycled over the trials)
PREFER_CONTROLS_ON_WARP_BITS = os.environ.get("QB_CTRL_WARP", "1") != "0"  # A/B switch, see DESIGN.md
# From this many (padded) qubits on a state sweep costs real HBM time, so a sweep whose FP64 work fits under its memory time is
# free: spread the arithmetic evenly over the sweeps (without adding a sweep) instead of packing the first ones full.
BALANCE_MIN_QUBITS = int(os.environ.get("QB_BALANCE_MIN_QUBITS", "22"))
BALANCE_SLACK = (1.0, 1.1, 1.25, 1.5)

# position kinds in the encoded program
K_NONE, K_REG, K_THREAD, K_EXT = 0, 1, 2, 3


@dataclass
class PassOp:
    op_index: int
    kind: int
    tgt_kind: int
    tgt_pos: int
    ctrl_kind: int
    ctrl_pos: int


@dataclass
class PassPlan:
    reg_bits: list[int] = field(default_factory=list)  # tile-local positions held in registers
    ops: list[PassOp] = field(default_factory=list)
    thread_bits: list[int] = field(default_factory=list)  # tile-local position of thread-index bit i
    warp_local_exchange: bool = False  # the exchange AFTER this pass stays inside each warp (no CTA barrier needed)


@dataclass
class SweepPlan:
    tile_qubits: list[int]
    passes: list[PassPlan]

    @property
    def n_ops(self) -> int:
        return sum(len(p.ops) for p in self.passes)


@dataclass
class CircuitPlan:
    n_qubits: int  # logical
    n_eff: int  # padded to >= tile_bits
    tile_bits: int
    reg_bits: int
    low_bits: int
    sweeps: list[SweepPlan]
    n_ops: int
    init_ops: list[int] = field(default_factory=list)  # per (padded) qubit: op giving its initial state, or -1

    @property
    def n_passes(self) -> int:
        return sum(len(s.passes) for s in self.sweeps)


def _select_sweep(ops: Sequence[KernelOp], remaining: list[int], n_eff: int, k: int, low: int, max_ops: int, rng=None, p_accept: float = 1.0,
                  max_cost: float = float("inf")):
    """Greedy choice of one sweep: walk the remaining ops in circuit order, take every op that is not blocked by a deferred one
    and whose target is (or can still become) a tile qubit.  With ``rng`` a new tile qubit is only accepted with probability
    ``p_accept`` -- the randomised restarts of ``plan_circuit`` use that to leave room for qubits with more work behind them.
    ``max_cost``: stop taking ops once the sweep's FP64 work (multiply-adds per amplitude) reaches it (sweep balancing)."""
    tile = set(range(min(low, n_eff)))
    pend_dense: set[int] = set()
    pend_any: set[int] = set()
    chosen: list[int] = []
    rest: list[int] = []
    cost = 0.0
    for i in remaining:
        op = ops[i]
        t, c = op.target, op.control
        dense = op.kind == DENSE
        blocked = (t in pend_any) if dense else (t in pend_dense)
        if c >= 0 and c in pend_dense:
            blocked = True
        ok = not blocked and len(chosen) < max_ops and cost < max_cost
        if ok and dense and t not in tile:
            if len(tile) < k and (rng is None or rng.random() < p_accept):
                tile.add(t)
            else:
                ok = False
        if ok:
            chosen.append(i)
            cost += dfma_per_amplitude(op)
        else:
            rest.append(i)
            if dense:
                pend_dense.add(t)
            pend_any.add(t)
            if c >= 0:
                pend_any.add(c)
    q = 0
    while len(tile) < min(k, n_eff):
        if q not in tile:
            tile.add(q)
        q += 1
    return sorted(tile), chosen, rest


def _thread_bit_order(reg: list[int], k: int, low: int, warp_pos: Sequence[int] = (), edge: bool = False) -> list[int]:
    """Which tile-local bit each thread-index bit carries: first the 5 lane bits, then the warp-index bits.

    * Passes that touch global memory (no low bit in registers) keep the ``low`` lowest tile bits on lanes 0.. so lanes
      walk contiguous amplitudes.  Pure shared-memory passes pick lane bits 0-2 from three different
      (position mod 3) classes when possible: the shared-memory layout XOR-folds the index in groups of three bits
      (qb_kernels.cuh: swz), which keeps the 128-bit accesses of a quarter warp conflict-free.
    * ``warp_pos``: tile positions the sweep wants on the warp-index bits (the highest thread-index bits).  When two
      consecutive passes carry the same positions there, every warp reads back exactly the amplitudes it wrote and the
      exchange between them needs ``__syncwarp`` only."""
    free = [b for b in range(k) if b not in reg]
    n_warp = max(0, len(free) - 5)
    warp = [b for b in warp_pos if b in free][:n_warp]
    rest = [b for b in free if b not in warp]
    while len(warp) < n_warp:  # not enough preferred positions available in this pass: take the highest free ones
        warp.insert(0, rest.pop())
    if not any(b < low for b in reg) or edge:
        return rest + warp  # ascending: the low bits that are not in registers sit on the lowest lanes
    picked: list[int] = []
    for cls in range(3):
        for b in rest:
            if b % 3 == cls and b not in picked:
                picked.append(b)
                break
    return picked + [b for b in rest if b not in picked] + warp


def _plan_passes(ops: Sequence[KernelOp], chosen: list[int], tile_qubits: list[int], r: int, low: int) -> list[PassPlan]:
    k = len(tile_qubits)
    pos = {q: i for i, q in enumerate(tile_qubits)}
    regs: list[list[int]] = [[]]
    allow_low: list[bool] = [ALLOW_LOW_EDGE_PASSES]
    members: list[list[int]] = [[]]
    last_dense: dict[int, int] = {}
    last_any: dict[int, int] = {}

    def new_pass(low_ok: bool):
        regs.append([])
        allow_low.append(low_ok)
        members.append([])

    for i in chosen:
        op = ops[i]
        t, c = op.target, op.control
        dense = op.kind == DENSE
        lo = last_any.get(t, 0) if dense else last_dense.get(t, 0)
        if c >= 0:
            lo = max(lo, last_dense.get(c, 0))
        p = lo
        if dense:
            tp = pos[t]
            while True:
                if p == len(regs):
                    new_pass(True)
                if tp in regs[p] or (len(regs[p]) < r and (allow_low[p] or tp >= low)):
                    break
                p += 1
            if tp not in regs[p]:
                regs[p].append(tp)
            last_dense[t] = p
            last_any[t] = p
        else:
            if p == len(regs):
                new_pass(True)
            last_any[t] = max(last_any.get(t, 0), p)
        if c >= 0:
            last_any[c] = max(last_any.get(c, 0), p)
        members[p].append(i)

    # the last pass stores straight to global memory: its register bits must avoid the low (lane) bits
    if any(b < low for b in regs[-1]) and not ALLOW_LOW_EDGE_PASSES:
        new_pass(False)
    # an empty leading pass is only needed when the next one holds low bits in registers
    if len(regs) > 1 and not members[0] and not any(b < low for b in regs[1]):
        regs.pop(0), allow_low.pop(0), members.pop(0)

    # pad every register set to r positions (positions nobody targets, highest first)
    padded: list[list[int]] = []
    for idx, reg in enumerate(regs):
        reg = list(reg)
        for cand in range(k - 1, -1, -1):
            if len(reg) >= min(r, k):
                break
            if cand not in reg and (cand >= low or allow_low[idx]):
                reg.append(cand)
        padded.append(sorted(reg))
    # Tile positions carried by the warp-index bits, pass by pass: keep the previous pass's choice whenever those
    # positions are still outside the register set (then the exchange between the two passes stays inside each warp),
    # otherwise prefer positions that stay free for the longest run of following passes.
    n_warp = max(0, k - r - 5)

    def warp_candidates(idx: int) -> list[int]:
        # lanes 0..low-1 must carry the low tile bits in every pass that touches global memory
        direct = not any(b < low for b in padded[idx]) or (ALLOW_LOW_EDGE_PASSES and idx in (0, len(padded) - 1))
        return [b for b in range(k) if b not in padded[idx] and not (direct and b < low)]

    def control_positions(idx: int) -> dict[int, int]:
        """tile positions (outside the pass's register set) that control ops of pass idx -> number of such ops"""
        count: dict[int, int] = {}
        for i in members[idx]:
            c = ops[i].control
            if c >= 0 and c in pos and pos[c] not in padded[idx]:
                count[pos[c]] = count.get(pos[c], 0) + 1
        return count

    warp_choice: list[list[int]] = []
    for idx in range(len(padded)):
        cand = warp_candidates(idx)
        # a control on a lane bit makes the warp run the whole gate with half its lanes masked off; on a warp-index bit
        # the warps whose control bit is 0 skip the gate: controls go first
        ctrl = control_positions(idx) if PREFER_CONTROLS_ON_WARP_BITS else {}
        first = sorted((b for b in cand if b in ctrl), key=lambda b: (-ctrl[b], -b))[:n_warp]
        cand = [b for b in cand if b not in first]
        keep = first + [b for b in (warp_choice[-1] if warp_choice else []) if b in cand]

        def free_run(b: int) -> int:
            run = 0
            for nxt in range(idx + 1, len(padded)):
                if b not in warp_candidates(nxt):
                    break
                run += 1
            return run

        others = sorted((b for b in cand if b not in keep), key=lambda b: (-free_run(b), -b))
        warp_choice.append(sorted((keep + others)[:n_warp]))

    passes: list[PassPlan] = []
    for idx, (reg, mem, warp_pos) in enumerate(zip(padded, members, warp_choice)):
        edge = ALLOW_LOW_EDGE_PASSES and idx in (0, len(padded) - 1)
        plan = PassPlan(reg_bits=reg, thread_bits=_thread_bit_order(reg, k, low, warp_pos, edge))
        for i in mem:
            op = ops[i]

            def locate(qubit: int):
                if qubit not in pos:
                    return K_EXT, qubit
                tp = pos[qubit]
                return (K_REG, reg.index(tp)) if tp in reg else (K_THREAD, tp)

            tk, tpos = locate(op.target)
            if op.kind == DENSE:
                assert tk == K_REG
            ck, cpos = (K_NONE, 0) if op.control < 0 else locate(op.control)
            plan.ops.append(PassOp(i, op.kind, tk, tpos, ck, cpos))
        passes.append(plan)
    for a, b in zip(passes, passes[1:]):
        a.warp_local_exchange = n_warp > 0 and a.thread_bits[5:] == b.thread_bits[5:]
    return passes


def split_product_prefix(ops: Sequence[KernelOp], n_qubits: int) -> tuple[list[int], list[int]]:
    """Peel the product-state prefix off a circuit that starts in |0...0>.

    Returns ``(init_ops, remaining)``: ``init_ops[q]`` is the index of an uncontrolled op that is the first thing to
    happen to qubit q (its matrix's first column is qubit q's initial single-qubit state) or -1 (qubit starts in
    |0>); ``remaining`` are the indices of the ops still to be applied.  Controlled ops whose control qubit has not
    been touched yet act on control = |0> and are dropped altogether -- in an EVQE circuit that is every ``cu3`` of the
    first layer (circuit_layer.py:164-189: a control qubit carries no gate of its own in its layer)."""
    FRESH, PRODUCT, BUSY = 0, 1, 2
    status = [FRESH] * n_qubits
    init_ops = [-1] * n_qubits
    remaining: list[int] = []
    for i, op in enumerate(ops):
        t, c = op.target, op.control
        if c >= 0:
            if status[c] == FRESH:
                continue  # control is |0>: identity
            status[c] = status[t] = BUSY
            remaining.append(i)
        elif status[t] == FRESH:
            status[t] = PRODUCT
            init_ops[t] = i
        else:
            status[t] = BUSY
            remaining.append(i)
    return init_ops, remaining


def plan_circuit(
    ops: Sequence[KernelOp],
    n_qubits: int,
    tile_bits: int = TILE_BITS,
    reg_bits: int = REG_BITS,
    low_bits: int = LOW_BITS,
    product_prefix: bool = True,
) -> CircuitPlan:
    n_eff = max(n_qubits, tile_bits)
    if product_prefix:
        init_ops, remaining = split_product_prefix(ops, n_qubits)
    else:
        init_ops, remaining = [-1] * n_qubits, list(range(len(ops)))
    init_ops = init_ops + [-1] * (n_eff - n_qubits)

    def build(rng, p_accept: float = 1.0, caps: Sequence[float] = ()) -> list[SweepPlan]:
        todo, out = list(remaining), []
        while todo:
            max_ops = MAX_SWEEP_OPS
            cap = caps[len(out)] if len(out) < len(caps) else float("inf")
            while True:  # the kernel stages at most MAX_SWEEP_OPS matrices / MAX_SWEEP_PASSES pass records per sweep
                tile_qubits, chosen, rest = _select_sweep(ops, todo, n_eff, tile_bits, low_bits, max_ops, rng, p_accept, cap)
                if not chosen:  # an unlucky draw accepted nothing: plain greedy always makes progress
                    tile_qubits, chosen, rest = _select_sweep(ops, todo, n_eff, tile_bits, low_bits, max_ops, max_cost=cap)
                passes = _plan_passes(ops, chosen, tile_qubits, reg_bits, low_bits)
                if len(passes) <= MAX_SWEEP_PASSES or max_ops == 1:
                    break
                max_ops = max(1, max_ops // 2)
            todo = rest
            out.append(SweepPlan(tile_qubits, passes))
        return out

    def count_sweeps(rng, p_accept: float, limit: int) -> int:
        """Sweeps this draw needs; gives up (returns ``limit``) as soon as it cannot end below ``limit``."""
        todo, n = list(remaining), 0
        while todo:
            if n + 1 >= limit:
                return limit
            _, chosen, rest = _select_sweep(ops, todo, n_eff, tile_bits, low_bits, MAX_SWEEP_OPS, rng, p_accept)
            if not chosen:
                _, chosen, rest = _select_sweep(ops, todo, n_eff, tile_bits, low_bits, MAX_SWEEP_OPS)
            todo, n = rest, n + 1
        return n

    sweeps = build(None)
    # Randomised restarts of the tile choice (the greedy fills a tile with the first qubits it meets): fewer sweeps = less
    # HBM traffic for the same arithmetic.  Deterministic per circuit structure; only the sweep count is evaluated per
    # trial, the best draw is then planned in full and kept if it needs fewer (sweeps, passes).
    if PLAN_TRIALS > 0 and len(sweeps) > 1:
        import random
        import zlib

        seed0 = zlib.crc32(repr([(op.kind, op.target, op.control) for op in ops]).encode())
        best_n, best = len(sweeps), []  # draws that reach the smallest sweep count: (seed, p_accept)
        # every trial walks the remaining ops once per sweep: cap the search for very long circuits (EVQE individuals are far below)
        trials = min(PLAN_TRIALS, PLAN_VISIT_BUDGET // max(1, len(remaining) * len(sweeps)))
        for trial in range(trials):
            p_accept = PLAN_ACCEPT[trial % len(PLAN_ACCEPT)]
            # a draw is only interesting if it beats the best count, or ties it while full builds are still wanted
            n = count_sweeps(random.Random(seed0 + trial), p_accept, best_n + 1 if len(best) < PLAN_FULL_BUILDS else best_n)
            if n < best_n:
                best_n, best = n, []
            if n == best_n and len(best) < PLAN_FULL_BUILDS:
                best.append((seed0 + trial, p_accept))

        def cost(candidate: list[SweepPlan]) -> tuple[int, int]:
            return len(candidate), sum(len(sw.passes) for sw in candidate)

        # among the draws with the fewest sweeps the one with the fewest passes (= shared-memory exchanges) wins
        winner = None
        for seed, p_accept in best:
            alt = build(random.Random(seed), p_accept)
            if cost(alt) < cost(sweeps):
                sweeps, winner = alt, (seed, p_accept)
        # Sweep balancing (HBM-resident states only): the greedy packs the early sweeps full and leaves the last ones nearly
        # empty; a sweep costs max(HBM time, FP64 time), so the same number of sweeps with the arithmetic spread evenly is
        # never slower and is faster whenever a light sweep can hide work under its memory time.  The first sweep of a
        # product-state start only writes (half the HBM time): it gets half a share.
        if n_eff >= BALANCE_MIN_QUBITS and len(sweeps) > 1:

            def sweep_costs(candidate: list[SweepPlan]) -> list[float]:
                return [sum(dfma_per_amplitude(ops[po.op_index]) for ps in sw.passes for po in ps.ops) for sw in candidate]

            costs = sweep_costs(sweeps)
            weights = [0.5 if (product_prefix and i == 0) else 1.0 for i in range(len(sweeps))]
            total, wsum = sum(costs), sum(weights)
            for slack in BALANCE_SLACK:
                caps = [slack * total * w / wsum for w in weights]
                if max(c / w for c, w in zip(costs, weights)) <= 1.05 * slack * total / wsum:
                    break  # already at least this even
                alt = build(random.Random(winner[0]) if winner else None, winner[1] if winner else 1.0, caps)
                if len(alt) == len(sweeps):
                    sweeps = alt
                    break
    if not sweeps:  # empty circuit: one identity sweep so that |0...0> gets materialised
        tile_qubits = list(range(tile_bits))
        reg = list(range(tile_bits - reg_bits, tile_bits))
        sweeps.append(SweepPlan(tile_qubits, [PassPlan(reg_bits=reg, thread_bits=_thread_bit_order(reg, tile_bits, low_bits))]))
    return CircuitPlan(n_qubits, n_eff, tile_bits, reg_bits, low_bits, sweeps, len(ops), init_ops)


# -------------------------------------------------------------------------------------------------
# flat encoding handed to the C-ABI (layout documented in include/queasars_b200.h)
# -------------------------------------------------------------------------------------------------
SWEEP_DTYPE = np.dtype([("tile_qubits", np.int32, (16,)), ("pass_begin", np.int32), ("pass_end", np.int32), ("op_begin", np.int32), ("op_end", np.int32)], align=True)
PASS_DTYPE = np.dtype([("reg_bits", np.int32, (7,)), ("flags", np.int32), ("op_begin", np.int32), ("op_end", np.int32), ("thread_bits", np.uint8, (12,))], align=True)
PASSOP_DTYPE = np.dtype(
    [("op_index", np.int32), ("kind", np.uint8), ("tgt_kind", np.uint8), ("tgt_pos", np.uint8), ("ctrl_kind", np.uint8), ("ctrl_pos", np.uint8), ("variant", np.uint8), ("ctrl_qubit", np.uint8), ("tgt_qubit", np.uint8)],
    align=True,
)
ANGLE_DTYPE = np.dtype(
    [("slot", np.int32, (4,)), ("slot2", np.int32, (4,)), ("coeff", np.float64, (4,)), ("coeff2", np.float64, (4,)), ("const", np.float64, (4,)), ("kind", np.int32), ("pad", np.int32)],
    align=True,
)


def _predecode(po: PassOp, tile_qubits: Sequence[int]) -> tuple[int, int, int]:
    """(variant, ctrl_qubit, tgt_qubit) of include/queasars_b200.h: the kernel's jump index and the global qubit
    whose index bit predicates / selects for operands that live outside the registers."""

    def global_qubit(kind: int, pos: int) -> int:
        if kind == K_THREAD:
            return tile_qubits[pos]
        return pos if kind == K_EXT else 0xFF

    cb = po.ctrl_pos if po.ctrl_kind == K_REG else -1
    cq = global_qubit(po.ctrl_kind, po.ctrl_pos)
    if po.kind == DENSE:
        return 6 * po.tgt_pos + (cb + 1), cq, 0xFF
    if cb >= 0:
        return 40, cq, global_qubit(po.tgt_kind, po.tgt_pos)
    if po.tgt_kind == K_REG:
        return 33 + po.tgt_pos, cq, 0xFF
    return 32, cq, global_qubit(po.tgt_kind, po.tgt_pos)


def encode_plan(plan: CircuitPlan, ops: Sequence[KernelOp]):
    """-> (sweeps, passes, pass_ops, op_angles, init_ops) arrays."""
    assert plan.tile_bits <= 16 and plan.reg_bits in (3, REG_BITS)
    sweeps = np.zeros(len(plan.sweeps), dtype=SWEEP_DTYPE)
    passes = np.zeros(plan.n_passes, dtype=PASS_DTYPE)
    pass_ops = np.zeros(max(1, sum(s.n_ops for s in plan.sweeps)), dtype=PASSOP_DTYPE)
    pi = oi = 0
    for si, sw in enumerate(plan.sweeps):
        sweeps[si]["tile_qubits"][: len(sw.tile_qubits)] = sw.tile_qubits
        sweeps[si]["pass_begin"] = pi
        sweeps[si]["op_begin"] = oi
        for ps in sw.passes:
            passes[pi]["reg_bits"][: len(ps.reg_bits)] = ps.reg_bits
            passes[pi]["thread_bits"][: len(ps.thread_bits)] = ps.thread_bits
            passes[pi]["flags"] = PASS_FLAG_WARP_LOCAL if ps.warp_local_exchange else 0
            passes[pi]["op_begin"] = oi
            for po in ps.ops:
                rec = pass_ops[oi]
                rec["op_index"], rec["kind"] = po.op_index, po.kind
                rec["tgt_kind"], rec["tgt_pos"] = po.tgt_kind, po.tgt_pos
                rec["ctrl_kind"], rec["ctrl_pos"] = po.ctrl_kind, po.ctrl_pos
                rec["variant"], rec["ctrl_qubit"], rec["tgt_qubit"] = _predecode(po, sw.tile_qubits)
                oi += 1
            passes[pi]["op_end"] = oi
            pi += 1
        sweeps[si]["pass_end"] = pi
        sweeps[si]["op_end"] = oi
    angles = np.zeros(max(1, len(ops)), dtype=ANGLE_DTYPE)
    for i, op in enumerate(ops):
        for j, a in enumerate(op.angles):
            angles[i]["slot"][j], angles[i]["coeff"][j], angles[i]["const"][j] = a.slot, a.coeff, a.const
            angles[i]["slot2"][j], angles[i]["coeff2"][j] = a.slot2, a.coeff2
        angles[i]["kind"] = op.kind
    init_ops = np.asarray(plan.init_ops if plan.init_ops else [-1] * plan.n_eff, dtype=np.int32)
    return sweeps, passes, pass_ops, angles, init_ops
