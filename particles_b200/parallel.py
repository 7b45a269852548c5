"""Particle-sharded filters: N particles partitioned over the GPUs of one box, one process
per GPU (``torch.distributed``, backend nccl), SURVEY.md section 8(e).

Per step every rank runs the fused kernels on its own shard and the ranks exchange ONE
message: an all-gather of 8 doubles per rank -- (max, sum exp, sum exp^2) of the shard's
inferential and auxiliary log-weights.  Every rank then forms the same global
log-normaliser, ESS, logLt increment and resampling decision (``k_finish`` in
csrc/smcb_filter.cu merges the triples in rank order, so all ranks hold identical bits).
Resampling is per shard ("island" scheme): a shard resamples its own N/G particles from its
own normalised weights and restarts them at log-weight  LSE_shard(aux) - LSE_all(w) + log G,
i.e. the shard keeps its share of the total mass, which keeps the likelihood estimator
unbiased without moving particles.  This is a different (also consistent) estimator from the
reference's single global resampling; logLt parity is statistical, and G = 1 reduces to the
reference exactly.

The helpers at the bottom restate the merge / restart algebra on the host (NumPy); the gloo
tests use them to check the scheme itself on CPU with world_size 2.
"""
import numpy as np

from .core import _FusedEngine


class ShardedFilter(_FusedEngine):
    """One rank's shard of a fused filter.  ``n_local`` particles here, ``world * n_local``
    in total; Philox counters are offset by the global particle index, so the union of the
    shards draws the same numbers as one big filter would."""

    def __init__(self, spec, n_local, scheme, ESSrmin, seed, rank, world, group=None, noise=None):
        if n_local % 2:
            raise ValueError("sharded filters need an even number of particles per rank")
        super().__init__(spec, n_local, scheme, ESSrmin, seed, noise=noise,
                         n_global=n_local * world, index_offset=rank * n_local,
                         world=world, rank=rank, group=group)


# ---------------------------------------------------------------------------
# host restatement of the exchange algebra (used by tests and by post-processing)
# ---------------------------------------------------------------------------
def merge_lse3(triples):
    """Merge per-shard (m, s, q) = (max, sum exp(v - m), sum exp(2 (v - m))) in rank order."""
    M, S, Q = -np.inf, 0.0, 0.0
    for m, s, q in triples:
        if m == -np.inf:
            continue
        if M == -np.inf:
            M, S, Q = m, s, q
            continue
        new = max(M, m)
        ea, eb = np.exp(M - new), np.exp(m - new)
        M, S, Q = new, S * ea + s * eb, Q * ea * ea + q * eb * eb
    return M, S, Q


def global_stats(triples, n_global):
    """log_mean, ESS of the union of the shards (Weights.__init__, resampling.py:217-226)."""
    M, S, Q = merge_lse3(triples)
    return M + np.log(S / n_global), S * S / Q


def island_restart(local_aux, global_w, world):
    """Log-weight a shard's particles restart from after a per-shard resampling."""
    return (np.log(local_aux[1]) + local_aux[0]) - (np.log(global_w[1]) + global_w[0]) + np.log(world)


def lse3_of(v):
    v = np.asarray(v, dtype=np.float64)
    m = v.max()
    e = np.exp(v - m)
    return m, e.sum(), (e * e).sum()
