"""Pauli-sum operator container with the subset of ``qiskit.quantum_info.SparsePauliOp``'s API the
reference uses (construction from labels, scalar arithmetic, ``compose``, ``sum``, ``to_list``,
``num_qubits``), plus the mask form the kernels consume.

Labels are little-endian like upstream: the right-most character acts on qubit 0
(/root/reference/queasars/utility/pauli_strings.py:38-41).  A term is stored as
``(x_mask, z_mask, coeff)`` with ``P = i^{popcount(x & z)} * X^x Z^z`` i.e. a 'Y' is ``x & z`` set.
"""
from __future__ import annotations

import numbers
from typing import Iterable, Sequence

import numpy as np


def _label_to_masks(label: str) -> tuple[int, int]:
    x = z = 0
    n = len(label)
    for pos, ch in enumerate(label):
        bit = 1 << (n - 1 - pos)
        if ch == "X":
            x |= bit
        elif ch == "Z":
            z |= bit
        elif ch == "Y":
            x |= bit
            z |= bit
        elif ch != "I":
            raise ValueError(f"invalid Pauli label character {ch!r}")
    return x, z


def _masks_to_label(x: int, z: int, n: int) -> str:
    chars = []
    for q in range(n - 1, -1, -1):
        xb, zb = (x >> q) & 1, (z >> q) & 1
        chars.append("IXZY"[xb + 2 * zb] if not (xb and zb) else "Y")
    return "".join(chars)


class SparsePauliOp:
    """Sum of weighted Pauli strings."""

    def __init__(self, data, coeffs=None, *, num_qubits: int | None = None):
        if isinstance(data, SparsePauliOp):
            self._n, self._x, self._z, self._c = data._n, list(data._x), list(data._z), np.array(data._c, dtype=complex)
            return
        labels = [data] if isinstance(data, str) else list(data)
        if not labels and num_qubits is None:
            raise ValueError("empty operator needs num_qubits")
        self._n = len(labels[0]) if labels else int(num_qubits)
        self._x, self._z = [], []
        for lab in labels:
            if len(lab) != self._n:
                raise ValueError("all Pauli labels must have the same length")
            x, z = _label_to_masks(lab)
            self._x.append(x)
            self._z.append(z)
        if coeffs is None:
            self._c = np.ones(len(labels), dtype=complex)
        else:
            self._c = np.array([complex(c) for c in np.atleast_1d(coeffs)], dtype=complex)
            if len(self._c) != len(labels):
                raise ValueError("coeffs length does not match the number of labels")

    # ------------------------------------------------------------------ constructors / views
    @classmethod
    def _raw(cls, n, xs, zs, cs):
        op = cls.__new__(cls)
        # coefficients live in ONE ndarray that ``coeffs`` hands out without copying (like qiskit's SparsePauliOp): in-place edits are
        # visible to the primitives' operator fingerprint at the cost of hashing a few kilobytes
        op._n, op._x, op._z, op._c = n, list(xs), list(zs), np.array([complex(c) for c in cs], dtype=complex)
        return op

    @classmethod
    def from_list(cls, obj: Iterable[tuple[str, complex]], num_qubits: int | None = None):
        obj = list(obj)
        return cls([lab for lab, _ in obj], [c for _, c in obj], num_qubits=num_qubits)

    @classmethod
    def from_sparse_list(cls, obj: Iterable[tuple[str, Sequence[int], complex]], num_qubits: int):
        labels, coeffs = [], []
        for paulis, qubits, coeff in obj:
            chars = ["I"] * num_qubits
            for ch, q in zip(paulis, qubits):
                chars[num_qubits - 1 - q] = ch
            labels.append("".join(chars))
            coeffs.append(coeff)
        return cls(labels, coeffs, num_qubits=num_qubits)

    @property
    def num_qubits(self) -> int:
        return self._n

    @property
    def coeffs(self) -> np.ndarray:
        return np.asarray(self._c, dtype=complex)

    @property
    def size(self) -> int:
        return len(self._c)

    def __len__(self):
        return len(self._c)

    def to_list(self) -> list[tuple[str, complex]]:
        return [(_masks_to_label(x, z, self._n), c) for x, z, c in zip(self._x, self._z, self._c)]

    def masks(self) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
        return (
            np.asarray(self._x, dtype=np.uint64),
            np.asarray(self._z, dtype=np.uint64),
            np.asarray(self._c, dtype=complex),
        )

    def is_diagonal(self) -> bool:
        return all(x == 0 for x in self._x)

    # ------------------------------------------------------------------ arithmetic
    def _check(self, other):
        if not isinstance(other, SparsePauliOp):
            return False
        if other._n != self._n:
            raise ValueError("operators act on different numbers of qubits")
        return True

    def __add__(self, other):
        if not self._check(other):
            return NotImplemented
        return SparsePauliOp._raw(self._n, self._x + other._x, self._z + other._z, np.concatenate([self._c, other._c]))

    def __radd__(self, other):
        if other == 0:  # allows builtin sum()
            return self
        return NotImplemented

    def __neg__(self):
        return SparsePauliOp._raw(self._n, self._x, self._z, [-c for c in self._c])

    def __sub__(self, other):
        if not self._check(other):
            return NotImplemented
        return self + (-other)

    def __mul__(self, other):
        if not isinstance(other, numbers.Number):
            return NotImplemented
        return SparsePauliOp._raw(self._n, self._x, self._z, [c * other for c in self._c])

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self * (1.0 / other)

    def compose(self, other: "SparsePauliOp", qargs=None, front: bool = False) -> "SparsePauliOp":
        """Operator product ``other @ self`` (upstream ``compose`` semantics; ``front`` swaps the order)."""
        self._check(other)
        a, b = (other, self) if not front else (self, other)  # result = a . b
        xs, zs, cs = [], [], []
        for xa, za, ca in zip(a._x, a._z, a._c):
            ya = bin(xa & za).count("1")
            for xb, zb, cb in zip(b._x, b._z, b._c):
                yb = bin(xb & zb).count("1")
                x, z = xa ^ xb, za ^ zb
                # (i^ya X^xa Z^za)(i^yb X^xb Z^zb) = i^(ya+yb) (-1)^{za.xb} X^x Z^z ; X^x Z^z = i^{-y} P
                y = bin(x & z).count("1")
                phase = (1j) ** ((ya + yb - y) % 4) * (-1) ** (bin(za & xb).count("1") & 1)
                xs.append(x)
                zs.append(z)
                cs.append(ca * cb * phase)
        return SparsePauliOp._raw(self._n, xs, zs, cs)

    dot = compose

    @staticmethod
    def sum(ops: Sequence["SparsePauliOp"]) -> "SparsePauliOp":
        ops = list(ops)
        if not ops:
            raise ValueError("Input list is empty")
        out = ops[0]
        for op in ops[1:]:
            out = out + op
        return out

    def simplify(self, atol: float = 1e-8) -> "SparsePauliOp":
        merged: dict = {}
        for x, z, c in zip(self._x, self._z, self._c):
            merged[(x, z)] = merged.get((x, z), 0.0) + c
        keep = [(k, c) for k, c in merged.items() if abs(c) > atol]
        if not keep:
            keep = [((0, 0), 0.0)]
        return SparsePauliOp._raw(self._n, [k[0] for k, _ in keep], [k[1] for k, _ in keep], [c for _, c in keep])

    def __repr__(self):
        return f"SparsePauliOp({[lab for lab, _ in self.to_list()]}, coeffs={self._c})"
