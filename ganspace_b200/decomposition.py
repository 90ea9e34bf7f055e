"""Decomposition driver: ``get_or_compute`` / ``compute`` -- the hot path's public entry point.

Mirror of /root/reference/decomposition.py (get_or_compute :362-368, _compute :370-402, compute :150-358,
linreg_lstsq :77-139, get_random_dirs :42-46, get_max_batch_size :49-74): same signature, validation
errors, cache-file naming, seeding protocol and 8-array ``.npz`` schema, so ``visualize.py``,
``interactive.py`` and the notebooks consume the result unchanged.

What changed is where the work happens.  The reference samples latents on the host, round-trips every
batch through host memory and runs sklearn's IncrementalPCA on the CPU; here
  * every NumPy-legacy latent stream of the run is generated on the GPU (one CTA per sample_latent seed),
  * the mapping network / hooked layer runs in hand-written CUDA kernels,
  * the IncrementalPCA merge chain runs on the device from per-group (mean, centred Gram) statistics,
  * the regression pass accumulates normal equations on the device,
so activations never leave HBM; the host only draws the seeds (NumPy global state, as the reference
does) and receives the final components.

Multi-GPU (one process per GPU, torch.distributed initialised): partial_fit group k belongs to rank k mod world; the run
proceeds in rounds of world x g groups: every rank computes the statistics of its g groups, one all-gather per round hands
them to every rank, and every rank merges them into its replica of the chain in the reference's order while the next round
is already being computed -- the result does not depend on the world size.
"""
from __future__ import annotations

import datetime
import inspect
import os
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from . import _native
from . import plan as _plan
from .config import Config  # noqa: F401  (re-exported like the reference module does)
from .estimators import get_estimator
from .models import get_instrumented_model
from .netdissect.nethook import InstrumentedModel

SEED_SAMPLING = 1
SEED_RANDOM_DIRS = 2
SEED_LINREG = 3
SEED_VISUALIZATION = 5

B = 20
n_clusters = 500

# bytes of latents generated per pipeline chunk (HBM is 180 GB; 8 GiB keeps config 2 in one chunk)
LATENT_CHUNK_BYTES = 8 << 30
# partial_fit groups whose statistics are computed by one set of launches (the first block is small so that the merge
# chain starts early; 10 groups x 10 tile pairs fill the 148 SMs once)
STATS_FIRST_BLOCK = int(os.environ.get("GANSPACE_B200_STATS_FIRST", 4))
STATS_BLOCK = int(os.environ.get("GANSPACE_B200_STATS_BLOCK", 10))


def get_random_dirs(components, dimensions):
    gen = np.random.RandomState(seed=SEED_RANDOM_DIRS)
    dirs = gen.normal(size=(components, dimensions))
    dirs /= np.sqrt(np.sum(dirs ** 2, axis=1, keepdims=True))
    return dirs.astype(np.float32)


def get_max_batch_size(inst, device, layer_name=None):
    """Largest probe batch (<= 20) whose peak memory stays under half the device (reference :49-74)."""
    inst.remove_edits()
    torch.cuda.reset_peak_memory_stats(device)
    total_mem = torch.cuda.get_device_properties(device).total_memory
    B_max = 20
    for i in range(2, B_max, 2):
        z = inst.model.sample_latent(n_samples=i)
        if layer_name:
            inst.model.partial_forward(z, layer_name)
        else:
            inst.model.forward(z)
        maxmem = torch.cuda.max_memory_allocated(device)
        del z
        if maxmem > 0.5 * total_mem:
            print("Batch size {:d}: memory usage {:.0f}MB".format(i, maxmem / 1e6))
            return i
    return B_max


def _dist():
    """(rank, world, group-is-live) of the data-parallel job, (0, 1, False) when not distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size(), True
    return 0, 1, False


class _StatsExchange:
    """All-gather of one round's per-group statistics (plan.rounds).  ``start`` enqueues the
    collective asynchronously; ``finish`` (called one round later, so that the next round's kernels are already queued behind
    it) merges the round's groups into this rank's chain replica in the reference's group order."""

    def __init__(self, d, world):
        self.d, self.world = d, world

    def start(self, rnd, means, grams):
        import torch.distributed as dist
        g = means.shape[0]
        recv_m = torch.empty((self.world * g, self.d), dtype=torch.float64, device=means.device)
        recv_g = torch.empty((self.world * g, self.d, self.d), dtype=torch.float64, device=means.device)
        works = [dist.all_gather_into_tensor(recv_g, grams, async_op=True),
                 dist.all_gather_into_tensor(recv_m, means, async_op=True)]
        return rnd, g, recv_m, recv_g, works, (means, grams)          # the send buffers stay alive until finish()

    def finish(self, pending, transformer, NB):
        rnd, g, recv_m, recv_g, works, _keep = pending
        for w in works:
            w.wait()
        for kk in rnd:
            slot = (kk % self.world) * g + (kk - rnd[0]) // self.world
            if not transformer.fit_partial_stats(NB, recv_m[slot], recv_g[slot]):
                return False
        return True


def _draw_seeds(count):
    """``count`` successive ``np.random.randint(int32 max)`` draws on the global state, i.e. the seeds that
    ``count`` successive ``sample_latent()`` calls would consume (models/wrappers.py:168-169)."""
    hi = np.iinfo(np.int32).max
    return [int(np.random.randint(hi)) for _ in range(count)]


def _sample_batches(model, Bsz, seeds, out=None):
    """Rows of len(seeds) consecutive ``sample_latent(Bsz)`` calls, as one [len*Bsz, ...] device tensor."""
    if hasattr(model, "sample_latents_multi"):
        return model.sample_latents_multi(Bsz, seeds, out=out)
    return torch.cat([model.sample_latent(Bsz, seed=s) for s in seeds], dim=0)


def _sample_batches_lazy(model, Bsz, seeds):
    """Like _sample_batches, plus ``ensure(row_end)``: rows [0, row_end) are final once it returns.  Models whose
    sample_latent has a second stage (StyleGAN2 W space: the mapping network) run it chunk by chunk on demand,
    so the IPCA chain starts on the first groups while later rows are still being mapped."""
    fn = getattr(model, "sample_latents_multi", None)
    if fn is not None and "lazy" in inspect.signature(fn).parameters:
        return fn(Bsz, seeds, lazy=True)
    return _sample_batches(model, Bsz, seeds), (lambda row_end: None)


# Solve for directions in latent space that match PCs in activation space (reference :77-139)
def linreg_lstsq(comp_np, mean_np, stdev_np, inst, config, affine=None, native=False):
    """``affine``: when the hooked layer is affine in the latent (models/biggan.py AffineLayer), comp/mean are
    given in its r-dimensional coordinates and the projections run there: (act-mean).comp^T == (y-ybar).comp_y^T.
    ``native``: comp/mean are given in the model's device feature order (``feature_layout``) and the activations
    come from ``model.activations_into`` in that same order (no hook read-back, no permutation)."""
    print("Performing least squares regression", flush=True)
    torch.manual_seed(SEED_LINREG)
    np.random.seed(SEED_LINREG)
    model = inst.model
    dev = model.device
    comp = torch.from_numpy(comp_np).float().to(dev).contiguous()
    mean = torch.from_numpy(mean_np).float().to(dev).reshape(-1).contiguous()
    stdev = torch.from_numpy(stdev_np).float().to(dev).contiguous()

    n_samp = max(10_000, config.n) // B * B
    n_comp = comp.shape[0]
    latent_dims = int(model.get_latent_dims())      # consumes one global draw, as in the reference (:88)
    rank, world, live = _dist()

    acc = _native.LinregAccumulator(n_comp, latent_dims, dev)
    act_buf = torch.empty((B, comp.shape[1]), dtype=torch.float32, device=dev) if native else None
    seeds = _draw_seeds(n_samp // B)
    group = max(1, min(len(seeds), (256 << 20) // max(1, B * latent_dims * 4)))
    for start in range(0, len(seeds), group):
        idx = [i for i in range(start, min(start + group, len(seeds))) if i % world == rank]
        if not idx:
            continue
        z_all = _sample_batches(model, B, [seeds[i] for i in idx])
        for j in range(len(idx)):
            z = z_all[j * B:(j + 1) * B]
            if affine is not None:
                act = affine.coords(z.reshape(B, -1))
            elif native:
                act = model.activations_into(z, config.layer, act_buf)
            else:
                with torch.no_grad():
                    model.partial_forward(z, config.layer)
                act = inst.retained_features()[config.layer].reshape(B, -1)
            acc.accumulate(act.contiguous(), comp, mean, stdev, z.reshape(B, -1).contiguous())
    if live:
        import torch.distributed as dist
        flat = acc.state.view(torch.float64)
        dist.all_reduce(flat)
    acc.n_total = n_samp
    M_t, Z_mean = acc.solve()
    return M_t.cpu().numpy()[:n_comp, :], Z_mean.cpu().numpy().reshape(1, -1)


def regression(comp, mean, stdev, inst, config, affine=None, native=False):
    M = np.dot(comp, comp.T)
    # fp32 components (large-d engine) are orthonormal to ~1e-6, the reference's float64 ones to 1e-15
    if not np.allclose(M, np.identity(M.shape[0]), atol=1e-8 if comp.dtype == np.float64 else 5e-6):
        det = np.linalg.det(M)
        print(f"WARNING: Computed basis is not orthonormal (determinant={det})")
    return linreg_lstsq(comp, mean, stdev, inst, config, affine=affine, native=native)


def compute(config, dump_name, instrumented_model):
    """decomposition.compute (:150-358): run the pipeline, rank 0 writes the 8-array .npz."""
    timestamp = lambda: datetime.datetime.now().strftime("%d.%m %H:%M")
    print(f"[{timestamp()}] Computing", dump_name.name)
    state = {}
    try:
        arrays = compute_arrays(config, instrumented_model, state)
    except _native.ChainNotConverged as err:
        # the iterated chain step found no spectral gap after component c (e.g. activations of numerical rank < n_components):
        # run again with sklearn's own exact per-step eigen-solve (every rank of a distributed run sees the same status)
        print(f"{err}\nRe-running with the direct chain step", flush=True)
        _native.set_chain_mode(True)
        try:
            state = {}
            arrays = compute_arrays(config, instrumented_model, state)
        finally:
            _native.set_chain_mode(False)
    if state.get("canceled_at") is not None:
        # Ctrl-C during the fitting loop: the reference saves what was fitted so far under n{gi} and exits 1 (:268-274,342-343)
        dump_name = dump_name.parent / dump_name.name.replace(f"n{state['N']}", f"n{state['canceled_at']}")
        print(f'Saving current state to "{dump_name.name}" before exiting')
    rank, world, live = _dist()
    if rank == 0:
        os.makedirs(dump_name.parent, exist_ok=True)
        # same 8-array .npz container, read by np.load exactly like the reference's; stored WITHOUT deflate: the arrays are
        # float32 noise (7 % smaller compressed) and single-core zlib costs 8-18 ms for config 2 -- a fifth of the whole
        # device run -- and minutes for conv feature maps (act_comp = 168 MB at convs.4).  GANSPACE_B200_NPZ_COMPRESS=1 restores
        # np.savez_compressed (decomposition.py:331-341).
        compress = os.environ.get("GANSPACE_B200_NPZ_COMPRESS") == "1" and sum(a.nbytes for a in arrays.values()) <= (64 << 20)
        (np.savez_compressed if compress else np.savez)(dump_name, **arrays)
    if live:
        import torch.distributed as dist
        dist.barrier()
    if state.get("canceled_at") is not None:
        sys.exit(1)


def _phase_timer():
    """GANSPACE_B200_TIMING=1: print wall-clock per phase (synchronising); otherwise a no-op."""
    if os.environ.get("GANSPACE_B200_TIMING") != "1":
        return lambda label: None
    import time
    state = {"t": time.perf_counter()}

    def tick(label):
        torch.cuda.synchronize()
        now = time.perf_counter()
        print(f"[timing] {label}: {now - state['t']:.3f} s", flush=True)
        state["t"] = now
    return tick


def compute_arrays(config, instrumented_model, state=None):
    """Everything of compute() up to (not including) the file write; returns the 8 float32 arrays.
    ``state`` (optional dict) receives N and, after a KeyboardInterrupt in the fitting loop, ``canceled_at`` = the
    number of samples fitted so far (single-process runs; the result then describes that prefix of the chain)."""
    global B
    tick = _phase_timer()
    state = {} if state is None else state

    torch.manual_seed(0)
    np.random.seed(0)

    device = _native.require_cuda("cuda")      # no CPU fallback (the reference falls back to 'cpu', :163-164)
    rank, world, live = _dist()
    layer_key = config.layer

    if instrumented_model is None:
        inst = get_instrumented_model(config.model, config.output_class, layer_key,
                                      torch.device("cuda", torch.cuda.current_device()))
        model = inst.model
    else:
        print("Reusing InstrumentedModel instance")
        inst = instrumented_model
        model = inst.model
        inst.remove_edits()
        model.set_output_class(config.output_class)
    device = model.device

    if config.use_w:
        print("Using W latent space")
        model.use_w()

    inst.retain_layer(layer_key)
    model.partial_forward(model.sample_latent(1), layer_key)
    sample_shape = inst.retained_features()[layer_key].shape
    sample_dims = int(np.prod(sample_shape))
    print("Feature shape:", sample_shape)

    input_shape = inst.model.get_latent_shape()
    input_dims = int(inst.model.get_latent_dims())

    config.components = min(config.components, sample_dims)
    # layers that are affine in the latent expose a thin factorisation act = (z R^T) Q^T + offset; the PCA then
    # runs on the r-dimensional coordinates and is lifted through the isometry Q at the end (models/biggan.py)
    affine = model.affine_layer(layer_key) if hasattr(model, "affine_layer") else None
    if affine is not None and config.components > affine.rank:
        raise NotImplementedError(f"components={config.components} exceeds the rank {affine.rank} of layer {layer_key}")
    transformer = get_estimator(config.estimator, config.components, config.sparsity, device=device)
    if not transformer.batch_support:
        raise RuntimeError("only batched estimators run on the device path")
    samples_are_latents = layer_key in ["g_mapping", "style"] and inst.model.latent_space_name() == "W"
    # conv feature maps (d up to ~10^6): the large-d IPCA engine keeps sklearn's stacked matrix in HBM and the model's
    # producer kernels write each batch into it in the device feature order (NHWC); the fixed NHWC->NCHW permutation
    # is applied once to the exported components (PCA is equivariant under it)
    large_d = affine is None and not samples_are_latents and sample_dims > transformer.transformer.SMALL_D_MAX
    layout = model.feature_layout(layer_key) if (large_d and hasattr(model, "feature_layout")) else None
    if large_d and live:
        # row-parallel generation + feature-sharded chain: rank r synthesises rows
        # [r NB/W, (r+1) NB/W) of every group, one all-to-all hands rank r the feature block r of all NB rows, the
        # small-side Gram is all-reduced, every rank solves the same small eigenproblem and updates its block.
        if layout is None:
            raise NotImplementedError("feature-sharded large-d runs need a model with activations_into / feature_layout")
        transformer.transformer.enable_feature_sharding(rank, world)
    if layout is not None:
        lh, lw, lc = layout[1]
        to_nchw = lambda A: np.ascontiguousarray(A.reshape(A.shape[0], lh, lw, lc).transpose(0, 3, 1, 2)).reshape(A.shape[0], -1)
        to_native = lambda A: np.ascontiguousarray(A.reshape(A.shape[0], lc, lh, lw).transpose(0, 2, 3, 1)).reshape(A.shape[0], -1)
    else:
        to_nchw = to_native = lambda A: A

    B = config.batch_size or get_max_batch_size(inst, device, layer_key)
    pl = _plan.make_plan(config.n, B, config.components)
    N, NB = pl.N, pl.NB
    print("B={}, N={}, dims={}, N/dims={:.1f}".format(B, N, sample_dims, N / sample_dims), flush=True)

    torch.manual_seed(config.seed or SEED_SAMPLING)
    np.random.seed(config.seed or SEED_SAMPLING)

    # ---- Phase A: the seeds of every sample_latent(B) call the reference makes (:232-236) ----------
    seeds = _draw_seeds(pl.n_calls)
    # W-space runs end with model.sample_latent(5000) for lat_stdev (:325-329).  Without a regression pass nothing touches the
    # global NumPy state in between, so its seed is the next draw; its latent stream (one sequential MT19937 stream, ~6 ms on
    # one SM) is generated on a side stream while the run proceeds instead of at the tail of the critical path.
    lat_stdev_z = None
    if config.use_w and samples_are_latents and hasattr(model, "draw_z_async"):
        lat_stdev_z = model.draw_z_async(5000, _draw_seeds(1)[0])

    # ---- Phase B: per-group statistics + merge chain (:239-265) ------------------------------------
    K = pl.K
    d = affine.rank if affine is not None else sample_dims
    groups_per_chunk = max(1, int(LATENT_CHUNK_BYTES // max(1, NB * input_dims * 4)))
    X = None
    tr = transformer.transformer
    state["N"] = N
    k = 0
    stop = False                 # fit_partial returned False (e.g. n_components > first batch): the reference leaves the loop (:262-263)
    exchange = _StatsExchange(d, world) if (live and not large_d) else None
    if not large_d and hasattr(tr, "begin_run"):
        tr.begin_run(K, NB, d, device)           # groups 1 .. K-1 merge inside one resident chain kernel (csrc/subspace.cu)

    def group_rows(rows):
        """[n, d] activations of the hooked layer for latent rows ``rows`` (small-d engine; n is a multiple of NB)."""
        if samples_are_latents:
            return rows
        if affine is not None:
            return affine.coords(rows)
        out = torch.empty((rows.shape[0], d), dtype=torch.float32, device=device)
        for g0 in range(0, rows.shape[0], NB):
            for mb in range(0, NB, B):
                z = rows[g0 + mb:g0 + mb + B].reshape(-1, *input_shape[1:])
                with torch.no_grad():
                    model.partial_forward(z, layer_key)
                batch = inst.retained_features()[layer_key].reshape((z.shape[0], -1))
                space_left = min(B, NB - mb)
                out[g0 + mb:g0 + mb + space_left] = batch[:space_left]
        return out

    try:
        for c0 in range(0, K, groups_per_chunk):
            if stop:
                break
            c1 = min(c0 + groups_per_chunk, K)
            if large_d:                              # every rank takes part in every group (its row range)
                mine = list(range(c0, c1))
            else:
                mine = _plan.groups_to_process(pl, rank, world, c0, c1)
            runs = _plan.contiguous_runs(mine)
            # every sample_latent call this rank needs for the chunk, generated by ONE launch (one CTA per seed)
            needed, offsets = _plan.batch_slots(pl, runs)
            lat, ensure_rows = _sample_batches_lazy(model, B, [seeds[b] for b in needed])
            lat = lat.reshape(lat.shape[0], -1)
            if large_d:
                for run, off in zip(runs, offsets):
                    if stop:
                        break
                    for k in run:
                        r = off + (k - run[0]) * NB
                        ensure_rows(r + NB)
                        rows = lat[r:r + NB]
                        X = tr.batch_buffer(NB, d, device)              # rows of the engine's stacked matrix, in HBM
                        lo, hi = (rank * (NB // world), (rank + 1) * (NB // world)) if live else (0, NB)   # this rank's rows
                        for mb in range(0, NB, B):
                            a, b_ = max(mb, lo), min(mb + min(B, NB - mb), hi)
                            if a >= b_:
                                continue
                            z = rows[a:b_].reshape(-1, *input_shape[1:])
                            if layout is not None:
                                model.activations_into(z, layer_key, X[a - lo:b_ - lo])
                            else:
                                with torch.no_grad():
                                    model.partial_forward(z, layer_key)
                                X[a - lo:b_ - lo] = inst.retained_features()[layer_key].reshape((z.shape[0], -1))
                        if not transformer.fit_partial_inplace(NB):
                            stop = True
                            break
            else:
                # small-d engine: rounds of world x g consecutive groups.  The statistics (mean, centred Gram) of this rank's
                # g groups of a round come out of one set of launches (tensor-core Gram, csrc/stats_tc.cu); the chain steps
                # follow in the reference's group order, on every rank, from the all-gathered statistics.
                where = {kk: off + (kk - run[0]) * NB for run, off in zip(runs, offsets) for kk in run}
                pending = None
                for rnd in _plan.rounds(c0, c1, world, STATS_FIRST_BLOCK if c0 == 0 else STATS_BLOCK, STATS_BLOCK):
                    if stop:
                        break
                    own = [kk for kk in rnd if _plan.owner(kk, world) == rank]
                    g_max = -(-len(rnd) // world)
                    means = torch.zeros((g_max, d), dtype=torch.float64, device=device)
                    grams = torch.zeros((g_max, d, d), dtype=torch.float64, device=device) if (live and len(own) < g_max) \
                        else torch.empty((g_max, d, d), dtype=torch.float64, device=device)
                    if own:
                        k = own[0]
                        contiguous = all(where[own[i]] == where[own[0]] + i * NB for i in range(len(own)))
                        spans = [own] if contiguous else [[kk] for kk in own]
                        i0 = 0
                        for span in spans:
                            r0 = where[span[0]]
                            ensure_rows(r0 + len(span) * NB)
                            Xb = group_rows(lat[r0:r0 + len(span) * NB])
                            _native.batch_stats_multi(Xb, len(span), NB, mean_out=means[i0:i0 + len(span)],
                                                      gram_out=grams[i0:i0 + len(span)])
                            i0 += len(span)
                        X = Xb[(len(span) - 1) * NB:]
                    if (K - 1) in rnd and _plan.owner(K - 1, world) != rank:
                        r0 = where[K - 1]                                # every rank keeps the final group's sample buffer
                        ensure_rows(r0 + NB)
                        X = group_rows(lat[r0:r0 + NB])
                    if exchange is None:
                        for i, kk in enumerate(own):
                            k = kk
                            if not transformer.fit_partial_stats(NB, means[i], grams[i]):
                                stop = True
                                break
                    else:
                        cur = exchange.start(rnd, means, grams)
                        if pending is not None and not exchange.finish(pending, transformer, NB):
                            stop = True
                        pending = cur
                        if rnd[0] < 2 * world * STATS_BLOCK and not stop:
                            # the first rounds are merged at once (the chain is idle and waiting for them); later rounds one
                            # round late, so that the collective overlaps the next round's kernels
                            if not exchange.finish(pending, transformer, NB):
                                stop = True
                            pending = None
                if pending is not None and not stop and not exchange.finish(pending, transformer, NB):
                    stop = True
            ensure_rows(lat.shape[0])
            del lat    # X (a view of the last group when samples_are_latents) keeps its storage alive
    except KeyboardInterrupt:
        if live:
            raise
        # the reference's `gi` of the interrupted group (:268-272) = the samples merged so far
        state["canceled_at"] = int(tr.n_samples_seen_)

    finally:
        if hasattr(tr, "end_run"):
            tr.end_run()                                # a resident chain kernel never waits for groups that will not come
    tick("sampling + activations + IPCA chain")
    # host work that does not depend on the chain's result, done while the device still runs the last merge steps (the
    # export below is the first call that waits for them): get_random_dirs' host stream + upload, the lat_stdev latents
    device_dirs = layout is not None and config.components * sample_dims >= (1 << 22)
    pre_dirs = None if device_dirs else \
        torch.from_numpy(to_native(get_random_dirs(config.components, int(np.prod(sample_shape))))).to(device, non_blocking=True)
    pre_lat = model.z_to_latent(lat_stdev_z()).reshape(5000, input_dims) if (config.use_w and lat_stdev_z is not None) else None
    X_comp, X_stdev, X_var_ratio = transformer.get_components()
    X_comp = np.array(X_comp, copy=True)
    mean_dev = tr.device_attributes()["mean"]
    if affine is None:
        X_global_mean = tr.mean_.reshape((1, sample_dims))
        Y_comp, Y_mean = X_comp, X_global_mean              # device feature order (== the reference's unless `layout`)
    else:
        # lift through the isometry: components = components_y Q^T (svd_flip's sign rule is applied on the lifted
        # rows, as sklearn would on the full activations), mean = mean_y Q^T + offset
        lifted = affine.lift_rows(torch.from_numpy(X_comp).to(device))
        idx = torch.argmax(lifted.abs(), dim=1)
        signs = torch.sign(lifted[torch.arange(lifted.shape[0], device=device), idx])
        Y_comp = X_comp * signs.cpu().numpy()[:, None]
        Y_mean = tr.mean_.reshape((1, d))
        X_comp = (lifted * signs[:, None]).cpu().numpy()
        X_global_mean = (affine.lift_rows(mean_dev[None, :]) + affine.offset[None, :]).cpu().numpy()

    assert X_comp.shape[1] == sample_dims and X_comp.shape[0] == config.components \
        and X_global_mean.shape[1] == sample_dims and X_stdev.shape[0] == config.components, "Invalid shape"

    if samples_are_latents:
        Z_comp = X_comp
        Z_global_mean = X_global_mean
    else:
        Z_comp, Z_global_mean = regression(Y_comp, Y_mean, X_stdev, inst, config, affine=affine, native=layout is not None)

    tick("export + regression")
    Z_comp /= np.linalg.norm(Z_comp, axis=-1, keepdims=True)

    # random projections of the last group's buffer, centred on the global mean (:289-291,312-316)
    n_rand_samples = min(5000, NB if large_d else X.shape[0])
    if device_dirs:
        # get_random_dirs' stream (RandomState(2).normal) drawn by the device generator: 42M normals at convs.4
        g = _native.legacy_normal([SEED_RANDOM_DIRS], config.components * sample_dims, device).view(config.components, -1)
        g = g / torch.linalg.vector_norm(g.double(), dim=1, keepdim=True).float()
        dirs_dev = g.view(config.components, lc, lh, lw).permute(0, 2, 3, 1).reshape(config.components, -1).contiguous()
    else:
        dirs_dev = pre_dirs
    if affine is not None:                                  # dirs . (x - mean) == (dirs Q) . (y - ybar)
        dirs_dev = _native.linear(dirs_dev, affine.Q.T.float().contiguous())
    sub = mean_dev
    if large_d:                                             # the engine centred the last group in place by its batch mean
        sub = mean_dev - tr.last_batch_mean()
        X = tr.last_batch_rows(n_rand_samples)              # (feature shards gathered when distributed)
    rand_std_dev = _native.project_std(X[:n_rand_samples], dirs_dev, sub=sub)       # read back below, with lat_stdev

    X_comp = to_nchw(X_comp).reshape(-1, *sample_shape)
    X_global_mean = to_nchw(X_global_mean).reshape(sample_shape)
    Z_comp = Z_comp.reshape(-1, *input_shape)
    Z_global_mean = Z_global_mean.reshape(input_shape)

    lat_stdev = np.ones_like(X_stdev)
    if config.use_w:
        if pre_lat is not None:
            samples = pre_lat
        else:
            samples = model.sample_latent(5000).reshape(5000, input_dims)
        zc = torch.from_numpy(Z_comp.reshape(-1, input_dims).astype(np.float32))
        both = torch.cat([rand_std_dev.reshape(-1), _native.project_std(samples.contiguous(), zc).reshape(-1)]).cpu().numpy()
        X_stdev_random, lat_stdev = both[:rand_std_dev.numel()], both[rand_std_dev.numel():]
    else:
        X_stdev_random = rand_std_dev.cpu().numpy()

    if hasattr(model, "check_numerics"):
        model.check_numerics()
    tick("random directions + layout")
    arrays = {
        "act_comp": X_comp.astype(np.float32),
        "act_mean": X_global_mean.astype(np.float32),
        "act_stdev": X_stdev.astype(np.float32),
        "lat_comp": Z_comp.astype(np.float32),
        "lat_mean": Z_global_mean.astype(np.float32),
        "lat_stdev": lat_stdev.astype(np.float32),
        "var_ratio": X_var_ratio.astype(np.float32),
        "random_stdevs": X_stdev_random.astype(np.float32),
    }
    if instrumented_model is None:
        inst.close()
        del inst
        del model
    return arrays


def get_or_compute(config, model=None, submit_config=None, force_recompute=False):
    if submit_config is None:
        wrkdir = str(Path(__file__).parent.resolve())
        submit_config = SimpleNamespace(run_dir_root=wrkdir, run_dir=wrkdir)
    return _compute(submit_config, config, model, force_recompute)


def _compute(submit_config, config, model=None, force_recompute=False):
    basedir = Path(submit_config.run_dir)

    if config.n is None:
        raise RuntimeError("Must specify number of samples with -n=XXX")
    if model and not isinstance(model, InstrumentedModel):
        raise RuntimeError('Passed model has to be wrapped in "InstrumentedModel"')
    if config.use_w and "StyleGAN" not in config.model:
        raise RuntimeError(f"Cannot change latent space of non-StyleGAN model {config.model}")

    transformer = get_estimator(config.estimator, config.components, config.sparsity)
    dump_name = "{}-{}_{}_{}_n{}{}{}.npz".format(
        config.model.lower(),
        config.output_class.replace(" ", "_"),
        config.layer.lower(),
        transformer.get_param_str(),
        config.n,
        "_w" if config.use_w else "",
        f"_seed{config.seed}" if config.seed else "",
    )
    dump_path = basedir / "cache" / "components" / dump_name

    if not dump_path.is_file() or force_recompute:
        print("Not cached")
        t_start = datetime.datetime.now()
        compute(config, dump_path, model)
        print("Total time:", datetime.datetime.now() - t_start)
    return dump_path
