"""NumPy interpreter of the encoded sweep program -- executes exactly the semantics the CUDA sweep kernel
implements (tile gather by ``tile_qubits``, per-pass register bits, REG/THREAD/EXT operand kinds), so the
planner + encoder can be validated on the CPU against the oracle.  Test infrastructure."""
import cmath
import math

import numpy as np

from queasars_b200.gate_list import DENSE, DIAG
from queasars_b200.schedule import K_EXT, K_NONE, K_REG, K_THREAD


def op_matrix(angle_rec, params):
    vals = [
        angle_rec["const"][j]
        + (angle_rec["coeff"][j] * params[angle_rec["slot"][j]] if angle_rec["slot"][j] >= 0 else 0.0)
        + (angle_rec["coeff2"][j] * params[angle_rec["slot2"][j]] if angle_rec["slot2"][j] >= 0 else 0.0)  # deferred phase of the previous gate
        for j in range(4)
    ]
    g, t, p, l = vals
    if angle_rec["kind"] == DIAG:
        return np.array([[cmath.exp(1j * g), 0], [0, cmath.exp(1j * (g + l))]])
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array(
        [[cmath.exp(1j * g) * c, -cmath.exp(1j * (g + l)) * s], [cmath.exp(1j * (g + p)) * s, cmath.exp(1j * (g + p + l)) * c]]
    )


def run_program(encoded, n_eff, tile_bits, params, state=None, reg_bits=4):
    sweeps, passes, pass_ops, angles, init_ops = encoded
    size = 1 << n_eff
    if state is None:
        state = np.ones(1, dtype=np.complex128)
        for q in range(n_eff):  # product-state start: qubit q = first column of op init_ops[q] (or |0>)
            v = np.array([1.0, 0.0], dtype=np.complex128) if init_ops[q] < 0 else op_matrix(angles[init_ops[q]], params)[:, 0]
            state = np.kron(v, state)
        used = {int(po["op_index"]) for po in pass_ops[: sum(int(p["op_end"] - p["op_begin"]) for p in passes)]}
        assert not used & {int(i) for i in init_ops if i >= 0}
    else:
        state = state.copy()
    e = np.arange(1 << tile_bits, dtype=np.int64)
    for sw in sweeps:
        tq = [int(q) for q in sw["tile_qubits"][:tile_bits]]
        tile_mask = sum(1 << q for q in tq)
        other = [q for q in range(n_eff) if not (tile_mask >> q) & 1]
        goff = np.zeros_like(e)
        for i, q in enumerate(tq):
            goff |= ((e >> i) & 1) << q
        for tau in range(1 << (n_eff - tile_bits)):
            base = 0
            for i, q in enumerate(other):
                base |= ((tau >> i) & 1) << q
            idx = base | goff
            tile = state[idx]
            for ps in passes[sw["pass_begin"] : sw["pass_end"]]:
                reg = [int(b) for b in ps["reg_bits"][:reg_bits]]
                assert len(set(reg)) == reg_bits and all(0 <= b < tile_bits for b in reg)
                for po in pass_ops[ps["op_begin"] : ps["op_end"]]:
                    m = op_matrix(angles[po["op_index"]], params)

                    def bit_of(kind, pos):
                        kind, pos = int(kind), int(pos)
                        if kind == K_REG:
                            return (e >> reg[pos]) & 1
                        if kind == K_THREAD:
                            assert pos not in reg
                            return (e >> pos) & 1
                        if kind == K_EXT:
                            assert not (tile_mask >> pos) & 1
                            return np.full_like(e, (base >> pos) & 1)
                        raise AssertionError(kind)

                    active = np.ones_like(e, dtype=bool) if po["ctrl_kind"] == K_NONE else bit_of(po["ctrl_kind"], po["ctrl_pos"]).astype(bool)
                    if po["kind"] == DIAG:
                        tb = bit_of(po["tgt_kind"], po["tgt_pos"])
                        factor = np.where(tb == 1, m[1, 1], m[0, 0])
                        tile = np.where(active, tile * factor, tile)
                    else:
                        assert po["tgt_kind"] == K_REG
                        tbit = reg[int(po["tgt_pos"])]
                        lo = e[((e >> tbit) & 1) == 0]
                        hi = lo | (1 << tbit)
                        x, y = tile[lo], tile[hi]
                        act = active[lo]
                        new = tile.copy()
                        new[lo] = np.where(act, m[0, 0] * x + m[0, 1] * y, x)
                        new[hi] = np.where(act, m[1, 0] * x + m[1, 1] * y, y)
                        tile = new
            state[idx] = tile
    return state
