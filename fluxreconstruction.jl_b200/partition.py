"""Element-slab partition of the structured meshes across ranks.

Pure host logic, shared by the GPU path and the gloo CPU tests: which rows/cells a rank
owns, who its neighbours are, and which rows travel each stage / each step.

2-D (euler2d): slabs along j (the second-fastest index, so each of the 4*nsp^2 planes
splits into contiguous chunks).  A rank's local array is [nx+2, nyl+2, nsp, nsp, 4]:
local row 0 / nyl+1 hold either the neighbour's boundary interior row of the *current*
stage (interior slab boundary) or the reference's frozen per-step ghost row (global
boundary, euler2d_wave.jl:127-132).
1-D: contiguous cell blocks with one halo cell on each side.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Slab:
    rank: int
    nranks: int
    n_global: int
    start: int   # first owned global index (1-based like the reference: 1..n_global)
    count: int   # owned rows / cells

    @property
    def stop(self):  # last owned global index (inclusive)
        return self.start + self.count - 1

    @property
    def lo(self):   # rank holding global index start-1 (periodic ring)
        return (self.rank - 1) % self.nranks

    @property
    def hi(self):
        return (self.rank + 1) % self.nranks

    @property
    def is_first(self):
        return self.rank == 0

    @property
    def is_last(self):
        return self.rank == self.nranks - 1


def slab(n_global: int, nranks: int, rank: int) -> Slab:
    """Balanced contiguous split: the first n_global % nranks ranks get one extra row."""
    if not (0 <= rank < nranks):
        raise ValueError("rank out of range")
    if n_global < nranks:
        raise ValueError("fewer rows than ranks")
    base, rem = divmod(n_global, nranks)
    count = base + (1 if rank < rem else 0)
    start = 1 + rank * base + min(rank, rem)
    return Slab(rank, nranks, n_global, start, count)


def stage_exchange_plan(s: Slab):
    """Messages of one RK stage for rank s: list of (peer, send_local_row, recv_local_row).
    Interior slab boundaries only -- the global seam is a frozen ghost row handled by
    ``step_ghost_plan``.  Local rows: 0 = lower halo, 1..count owned, count+1 = upper halo."""
    plan = []
    if not s.is_first:
        plan.append((s.lo, 1, 0))
    if not s.is_last:
        plan.append((s.hi, s.count, s.count + 1))
    return plan


def step_ghost_plan(s: Slab, periodic_y: bool):
    """Messages of the per-step ghost fill across the global y seam (periodic copy of
    euler2d_wave.jl:129-132): rank 0's lower ghost row <- last rank's top owned row, the last
    rank's upper ghost row <- rank 0's first owned row."""
    plan = []
    if not periodic_y or s.nranks == 1:
        return plan
    if s.is_first:
        plan.append((s.nranks - 1, 1, 0))
    if s.is_last:
        plan.append((0, s.count, s.count + 1))
    return plan
