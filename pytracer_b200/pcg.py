"""Host copy of the PCG32 XSH-RR generator (pcg.py:22-62 in the reference).

The host needs it only for bookkeeping: the renderer reads ``state``/``inc`` of the generator the
caller hands over and, after the launch, moves the object forward by the number of draws the
reference would have made (``advance`` — an O(log n) LCG jump), so that code following
``fire_all_rays`` sees the generator where the reference would have left it.  The per-ray draws
themselves happen in the CUDA library (csrc/rt_pcg.cuh).
"""
from __future__ import annotations

_MASK64 = (1 << 64) - 1
PCG_MULT = 6364136223846793005


class PCG:
    def __init__(self, init_state: int = 42, init_seq: int = 54):
        self.state = 0
        self.inc = ((init_seq << 1) | 1) & _MASK64
        self.random()
        self.state = (self.state + init_state) & _MASK64
        self.random()

    def random(self) -> int:
        old = self.state
        self.state = (old * PCG_MULT + self.inc) & _MASK64
        xorshifted = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF

    def random_float(self) -> float:
        return self.random() / 0xFFFFFFFF

    def advance(self, delta: int) -> None:
        """Jump ``delta`` draws ahead: compose x -> a*x + c with itself by repeated squaring."""
        self.state = lcg_advance(self.state, self.inc, delta)

    def __eq__(self, other) -> bool:
        return isinstance(other, PCG) and (self.state, self.inc) == (other.state, other.inc)

    def __repr__(self) -> str:
        return f"PCG(state={self.state}, inc={self.inc})"


def lcg_advance(state: int, inc: int, delta: int) -> int:
    acc_mult, acc_plus = 1, 0
    cur_mult, cur_plus = PCG_MULT, inc
    delta &= _MASK64
    while delta:
        if delta & 1:
            acc_mult = (acc_mult * cur_mult) & _MASK64
            acc_plus = (acc_plus * cur_mult + cur_plus) & _MASK64
        cur_plus = ((cur_mult + 1) * cur_plus) & _MASK64
        cur_mult = (cur_mult * cur_mult) & _MASK64
        delta >>= 1
    return (acc_mult * state + acc_plus) & _MASK64
