"""Host-side batch plan of decomposition.compute and its data-parallel sharding (pure Python, no device).

Restates the reference's sizing rules (decomposition.py:198-232) and defines how the partial_fit groups
are distributed over ranks: group k belongs to rank k mod world; the run proceeds in ROUNDS of
world x g consecutive groups; in a round every rank computes the statistics (mean, centred Gram) of its g groups with one
set of launches, one all-gather hands every rank all of the round's statistics, and every rank merges them into its replica
of the chain in the reference's group order -- while the next round's statistics are already being computed.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Sequence


@dataclass(frozen=True)
class Plan:
    B: int          # rows per sample_latent() call
    N: int          # samples entering the PCA  (n // B * B,               decomposition.py:201)
    NB: int         # rows per partial_fit      (max(B, 2000, 3c),         decomposition.py:220)
    n_lat: int      # latents drawn in phase A  (((N+NB-1)//B+1)*B,        decomposition.py:232)
    K: int          # number of partial_fit groups (range(0, N, NB),       decomposition.py:245)

    @property
    def n_calls(self) -> int:
        return self.n_lat // self.B

    def group_rows(self, k: int):
        return k * self.NB, k * self.NB + self.NB

    def batches_covering(self, row0: int, row1: int):
        """sample_latent call indices [b0, b1) whose rows cover [row0, row1)."""
        return row0 // self.B, (row1 + self.B - 1) // self.B


def make_plan(n: int, B: int, components: int) -> Plan:
    N = n // B * B
    NB = max(B, max(2_000, 3 * components))
    n_lat = ((N + NB - 1) // B + 1) * B
    K = len(range(0, N, NB))
    return Plan(B=B, N=N, NB=NB, n_lat=n_lat, K=K)


def owner(k: int, world: int, K: int = 0) -> int:
    """Round-robin ownership: the chain consumes the groups in order, so with k mod world every rank contributes to every
    round and the replicated chain can start on group 0 while later groups are still being produced."""
    return k % world


def groups_to_process(plan: Plan, rank: int, world: int, first: int = 0, last: int = None) -> List[int]:
    """Groups a rank touches in [first, last): the ones it owns, plus the final group on every rank (its
    sample buffer feeds random_stdevs, decomposition.py:313-316)."""
    last = plan.K if last is None else last
    return [k for k in range(first, last) if owner(k, world) == rank or k == plan.K - 1]


def rounds(first: int, last: int, world: int, per_rank_first: int, per_rank_later: int) -> List[range]:
    """Consecutive group ranges covering [first, last): the first round holds world*per_rank_first groups (a small first
    round lets the merge chain start early), the others world*per_rank_later."""
    out: List[range] = []
    a = first
    while a < last:
        size = world * (per_rank_first if a == first else per_rank_later)
        out.append(range(a, min(last, a + max(1, size))))
        a += max(1, size)
    return out


def contiguous_runs(ks: Sequence[int]) -> List[List[int]]:
    runs: List[List[int]] = []
    for k in ks:
        if runs and k == runs[-1][-1] + 1:
            runs[-1].append(k)
        else:
            runs.append([k])
    return runs


def slot_width(d: int) -> int:
    """doubles per statistics slot: centred Gram [d*d] followed by the batch mean [d]."""
    return d * d + d


def batch_slots(plan: Plan, runs: Sequence[Sequence[int]]):
    """sample_latent call indices needed by `runs` (sorted, unique) and, per run, the row offset of its first
    group inside the concatenated buffer of those calls."""
    needed: List[int] = []
    for run in runs:
        r0, r1 = plan.group_rows(run[0])[0], plan.group_rows(run[-1])[1]
        b0, b1 = plan.batches_covering(r0, r1)
        needed.extend(range(b0, b1))
    needed = sorted(set(needed))
    slot = {b: idx for idx, b in enumerate(needed)}
    offsets = []
    for run in runs:
        r0 = plan.group_rows(run[0])[0]
        b0 = r0 // plan.B
        offsets.append(slot[b0] * plan.B + (r0 - b0 * plan.B))
    return needed, offsets
