"""The ``test_dataset_to_add_noise`` path as one device pipeline (what ``bench.py`` times).

Per batch of clips (this rank's shard): bit-exact clip power -> SNR sigma -> white-noise mix fused
into the MFCC launch -> dataset standardisation of the feature rows (two-pass column statistics,
all-reduced over NCCL when sharded) -> standardised ``(B, n_mfcc*T)`` rows.  Mirrors the loop body
of ``Voice digit recogniton/attacks.py:402-407`` (``black_box_attack_on_audio_dataset_snr`` then
``standardize_dataset``).

Sigma is the reference's own chain (``sigma_mode="host"``, the default): the device computes ``P = mean(x**2)`` bit
for bit, its 4*B bytes go to pinned host memory (``frontend.copy_mapped``: a kernel, not a DMA copy), ``frontend.snr_sigma_host`` runs the four lines of
``attacks.py:235-241`` on it (numpy's log10, libm's powf) and sigma goes back - np.log10 / powf are not correctly
rounded and differ between hosts, so only the host can reproduce them.  To keep that round trip off the critical
path a caller that knows its next batch passes ``prefetch=``: the power launch and read-back of the NEXT step are
enqueued on a side stream right after THIS step's launches (``ASR_B200_POWER_STREAM=0``: on the caller's stream, in front
of this step's MFCC launch), so the host chain of step i+1 runs while the GPU is busy with step i and the power pass -
a streaming read of the audio - shares the device with the step's launches instead of standing in front of them.
``sigma_mode="device"`` evaluates the chain in float64 on the device (within 1 ulp of the host chain, not equal).

The launches of one step are short (tens of microseconds each) and their number is fixed, so the step is
captured once per batch buffers in CUDA graphs and replayed: a single graph for the whole step, on one GPU and when clips
are sharded over several GPUs - the ONE exchange of the path, the all-gather of the ranks' standardisation messages
(`Standardizer.exchange`), is then a peer-memory kernel of this library inside the graph (with the NCCL fallback transport:
two graphs with the collective between them).
"""
from __future__ import annotations

import os

from typing import Optional

import torch

from .frontend import (ClipBatch, MfccPlan, Noise, Standardizer, clip_power, snr_sigma_device, snr_sigma_host, randn,
                       babble_stream, babble_gain_host, copy_mapped)
from .params import MfccParams

# kernels of libasr_b200 launched by one `run_device` step besides the MFCC launches (`plan.launches`):
# standardisation: pass 1, pass 2, apply (single GPU; sharded: + message and merge); + power (and the sigma kernel in
# device mode, the babble stream with babble noise) when noisy; the e2e step adds randn
LAUNCHES_CMVN = 3
LAUNCHES_CMVN_SHARDED = 6       # pass 1, pass 2, message, peer-memory exchange, merge, apply
LAUNCHES_NOISE = 1              # the power pass; host sigma mode adds the two small copy kernels (power out, sigma in), device mode the sigma kernel


class _StepGraphs:
    """Captured launch groups of one step for fixed buffers."""

    def __init__(self):
        self.graphs = []       # replayed in order
        self.exchange_after = -1   # index of the graph followed by the (uncaptured) NCCL all-gather
        self.out = None
        self.keep = None       # tensors the graphs reference


class NoisyFeaturePipeline:
    def __init__(self, params: MfccParams, out_frames: int, device=None, distributed: bool = False, group=None,
                 world_size: int = 1, use_graphs: bool = True, path: str = "auto", sigma_mode: str = "host"):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.plan = MfccPlan(params, self.device.index, path=path)
        self.out_frames = int(out_frames)
        self.rows = self.plan.feature_rows
        self.D = self.rows * self.out_frames
        self.world_size = world_size
        self.distributed = distributed
        self.std = Standardizer(self.D, device=self.device, group=group, distributed=distributed)
        self.use_graphs = use_graphs
        # NCCL transport, opt-in (ASR_B200_CAPTURE_COLLECTIVES=1): capture the all-gather inside the step's graph.  Measured:
        # identical rows, no gain at 8 GPUs and the process group then takes minutes to tear down at exit.
        self.capture_collectives = os.environ.get("ASR_B200_CAPTURE_COLLECTIVES", "0") == "1"
        if self.std.exchange_transport == "p2p":
            self.capture_collectives = True     # the peer-memory exchange is one kernel of this library: the sharded step is ONE graph
        if sigma_mode not in ("host", "device"):
            raise ValueError("sigma_mode must be 'host' (the reference's chain, bit-exact) or 'device'")
        self.sigma_mode = sigma_mode
        self._sig = None             # host-sigma state: three pinned result slots, one device sigma vector
        # the power pass of a prefetched batch: 1 (default) = on a side stream, free to start as soon as the step has been
        # launched; 0 = on the caller's stream, in front of the step's MFCC launch.  Measured on the default step
        # (profiles/r2_power_overlap_ab.txt): 0.813 against 0.846 ms; releasing the pass only behind the step's MFCC launches
        # (two graphs with an event between them) 0.830; a 2-warp launch shaped to sit beside the persistent MFCC kernel > 1.5
        self._pow_mode = int(os.environ.get("ASR_B200_POWER_STREAM", "1"))
        if self._pow_mode not in (0, 1):
            raise ValueError("ASR_B200_POWER_STREAM must be 0 or 1")
        self._pow_stream = torch.cuda.Stream(self.device) if self._pow_mode != 0 else None
        self._pow_entry = torch.cuda.Event()
        self._pow_last = None        # event of the last side-stream power pass (join())
        self._feats = None
        self._cache: dict = {}
        self.ev_mfcc = None          # optional (start, end) CUDA events around the MFCC launch (eager steps only)

    def launches_per_step(self, noisy: bool, standardize: bool = True) -> int:
        noise = (LAUNCHES_NOISE + (1 if self.sigma_mode == "device" else 2)) if noisy else 0
        sharded = LAUNCHES_CMVN_SHARDED - (0 if self.std.exchange_transport == "p2p" else 1)   # the NCCL all-gather is not a kernel of this library
        cmvn = (sharded if self.distributed else LAUNCHES_CMVN) if standardize else 0
        return self.plan.launches(noisy) + noise + cmvn

    # ---- sigma: device power -> host chain -> device sigma, one or two steps ahead when the caller prefetches ----
    @staticmethod
    def _bkey(batch):
        return (batch.audio.data_ptr(), batch.n_clips, batch.max_length)

    def _sig_state(self, B: int):
        st = self._sig
        if st is None or st["cap"] < B:
            # three slots: the pass being consumed, and up to two prefetched ones (a caller may run two steps ahead)
            st = {"cap": B, "turn": 0, "pending": {},
                  "sigma_dev": torch.empty(B, dtype=torch.float64, device=self.device),
                  "slots": [{"P_dev": torch.empty(B, dtype=torch.float32, device=self.device),
                             "P_host": torch.empty(B, dtype=torch.float32).pin_memory(),
                             "sig_host": torch.empty(B, dtype=torch.float64).pin_memory(),
                             "event": torch.cuda.Event()} for _ in range(3)]}
            self._sig = st
            self._cache.clear()                            # captured graphs reference the old sigma vector
        return st

    def _submit_power(self, batch, babble: bool = False) -> int:
        """Enqueue power (+ the babble stream and its power) and the read-back of `batch` on the current stream; returns the slot."""
        st = self._sig_state(batch.n_clips)
        k = st["turn"]
        st["turn"] = (k + 1) % len(st["slots"])
        for key in list(st["pending"]):
            q = [v for v in st["pending"][key] if v != k]  # a prefetch that was never consumed loses its slot
            if q:
                st["pending"][key] = q
            else:
                del st["pending"][key]
        sl = st["slots"][k]
        B = batch.n_clips
        torch.cuda.current_stream(self.device).wait_event(sl["event"])   # an earlier pass into this slot (possibly on the other stream) is over
        clip_power(batch, out=sl["P_dev"])
        copy_mapped(sl["P_host"][:B], sl["P_dev"][:B])     # a kernel, not a DMA copy: never behind a bulk transfer of its direction
        if babble:
            n = batch.audio.shape[0]
            if sl.get("b_dev") is None or sl["b_dev"].numel() < n:
                sl["b_dev"] = torch.zeros(n, dtype=torch.float64, device=self.device)
                sl["Pb_dev"] = torch.empty(st["cap"], dtype=torch.float64, device=self.device)
                sl["Pb_host"] = torch.empty(st["cap"], dtype=torch.float64).pin_memory()
            babble_stream(batch, out=sl["b_dev"], power=sl["Pb_dev"])
            copy_mapped(sl["Pb_host"][:B], sl["Pb_dev"][:B])
        sl["event"].record()
        return k

    def prefetch_power(self, batch, babble: bool = False, after: Optional[torch.cuda.Event] = None) -> None:
        """Enqueue the power pass of a batch a later `run_device(..., snr_db)` will use (host sigma mode).  With the side
        stream the pass starts once `after` (an event of the caller's stream: everything the batch depends on) has
        completed and runs beside whatever the caller's stream holds from then on."""
        if not (self.sigma_mode == "host" or babble):
            return
        self._sig_state(batch.n_clips)
        key = self._bkey(batch) + (babble,)
        if self._pow_stream is None:
            k = self._submit_power(batch, babble)
            self._sig["pending"].setdefault(key, []).append(k)
            return
        if after is None:
            after = self._pow_entry
            after.record(torch.cuda.current_stream(self.device))
        self._pow_stream.wait_event(after)
        with torch.cuda.stream(self._pow_stream):
            k = self._submit_power(batch, babble)
        self._sig["pending"].setdefault(key, []).append(k)
        self._pow_last = self._sig["slots"][k]["event"]

    def join(self) -> None:
        """The caller's stream waits for the last prefetched power pass and for the uploads `run_host` has issued ahead
        (a timed region ends with them)."""
        cur = torch.cuda.current_stream(self.device)
        if self._pow_last is not None:
            cur.wait_event(self._pow_last)
        if getattr(self, "_hs", None) is not None:
            cur.wait_stream(self._s_h2d)

    def _sigma_for(self, batch, snr_db, prefetch, babble: bool = False):
        """(device sigma vector, noise stream or None) for this step, valid in stream order until the next call.  White
        noise: sigma of VDR/attacks.py:235-241, the stream is the caller's z.  Babble: the stream is the batch's babble
        sum and "sigma" the gain that gives it the white-noise sigma as RMS."""
        if self.sigma_mode == "device" and not babble:
            return snr_sigma_device(clip_power(batch), snr_db), None
        st = self._sig_state(batch.n_clips)
        key = self._bkey(batch) + (babble,)
        q = st["pending"].get(key)
        k = None
        if q:
            k = q.pop(0)                                    # the oldest prefetched pass of this batch
            if not q:
                del st["pending"][key]
        if k is None:
            k = self._submit_power(batch, babble)
        if prefetch is not None and self._pow_stream is None:
            self.prefetch_power(prefetch, babble)           # in front of this step's MFCC launch
            st = self._sig
        sl = st["slots"][k]
        B = batch.n_clips
        sl["event"].synchronize()
        sig = sl["sig_host"].numpy()[:B]
        snr_sigma_host(sl["P_host"].numpy()[:B], snr_db, out=sig)
        if babble:
            sig[:] = babble_gain_host(sig, sl["Pb_host"].numpy()[:B])
        copy_mapped(st["sigma_dev"][:B], sl["sig_host"][:B])
        return st["sigma_dev"][:B], (sl["b_dev"] if babble else None)

    def _feat_buffer(self, B: int) -> torch.Tensor:
        if self._feats is None or self._feats.shape[0] != B:
            self._feats = torch.empty((B, self.rows, self.out_frames), dtype=torch.float32, device=self.device)
        return self._feats

    # ---- the three launch groups of a step ----------------------------------------------------------
    def _group1(self, batch, z, snr_db, feats, sigma=None):
        noise = None
        if snr_db is not None:
            if sigma is None:
                sigma = snr_sigma_device(clip_power(batch), snr_db)
            noise = Noise.white(z, sigma)
        if self.ev_mfcc is not None:
            self.ev_mfcc[0].record()
        self.plan.mfcc(batch, out_frames=self.out_frames, noise=noise, out=feats)
        if self.ev_mfcc is not None:
            self.ev_mfcc[1].record()
        return noise

    def run_device(self, batch: ClipBatch, z: Optional[torch.Tensor], snr_db: Optional[float],
                   standardize: bool = True, out_dtype=torch.float32, prefetch: Optional[ClipBatch] = None,
                   noise_kind: str = "white") -> torch.Tensor:
        """Inputs resident in HBM; the launches are asynchronous on the current stream (in host sigma mode the call
        waits for the 4*B-byte power read-back of THIS batch, which a previous call's ``prefetch=`` has usually
        already enqueued).  ``prefetch``: the batch of the next noisy call.  ``noise_kind``: "white" (z is the caller's
        standard-normal stream) or "babble" (z is ignored: the noise is the sum of six other clips of the batch)."""
        sigma = None
        babble = noise_kind == "babble"
        side = snr_db is not None and (self.sigma_mode == "host" or babble) and prefetch is not None and self._pow_stream is not None
        if side:
            self._pow_entry.record(torch.cuda.current_stream(self.device))   # what the prefetched batch may depend on
        if snr_db is not None and (self.sigma_mode == "host" or babble):
            sigma, zb = self._sigma_for(batch, snr_db, prefetch, babble)
            if babble:
                z = zb
        if self.use_graphs and self.ev_mfcc is None:
            out = self._run_graphed(batch, z, snr_db, standardize, out_dtype, sigma)
        else:
            feats = self._feat_buffer(batch.n_clips)
            self._group1(batch, z, snr_db, feats, sigma)
            out = feats.view(batch.n_clips, self.D)
            if standardize:
                out = self.std.fit_transform(out, out_dtype=out_dtype)
        if side:
            self.prefetch_power(prefetch, babble, after=self._pow_entry)     # behind this step's launches, on the side stream
        return out

    def _run_graphed(self, batch, z, snr_db, standardize, out_dtype, sigma=None):
        # host sigma mode: the graph reads the (fixed) device sigma vector, so one capture serves every SNR
        key = (batch.audio.data_ptr(), batch.n_clips, batch.max_length, None if z is None else z.data_ptr(),
               snr_db if sigma is None else (snr_db is not None), standardize, out_dtype)
        sg = self._cache.get(key)
        if sg is None:
            if len(self._cache) > 32:
                self._cache.clear()
            sg = self._capture(batch, z, snr_db, standardize, out_dtype, sigma)
            self._cache[key] = sg
        for i, g in enumerate(sg.graphs):
            g.replay()
            if i == sg.exchange_after:
                self.std.exchange()                        # the one collective of the step (NCCL transport)
        return sg.out

    def _capture(self, batch, z, snr_db, standardize, out_dtype, sigma=None) -> _StepGraphs:
        B = batch.n_clips
        sg = _StepGraphs()
        feats = torch.empty((B, self.rows, self.out_frames), dtype=torch.float32, device=self.device)
        flat = feats.view(B, self.D)
        out = torch.empty((B, self.D), dtype=out_dtype, device=self.device) if standardize else flat
        sg.out, sg.keep = out, (batch, z, feats)
        # warm the kernels once outside capture (lazy module loading is not capturable)
        self._group1(batch, z, snr_db, feats, sigma)
        if standardize:
            self.std.fit_transform(flat, out=out)
        torch.cuda.synchronize(self.device)
        pool = torch.cuda.graph_pool_handle()

        def cap(fns):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                for fn in fns:
                    fn()
            sg.graphs.append(g)

        def mfcc():
            self._group1(batch, z, snr_db, feats, sigma)

        def message():                                      # sharded: this rank's statistics message
            self.std.local_stats([flat])
            self.std.local_message()

        def finish():                                       # sharded: after the all-gather
            self.std.merge()
            self.std.transform(flat, out=out)

        # segments of the step: one graph, or two around the exchange when the collective is an NCCL call that is not captured
        if not standardize:
            segs = [[mfcc]]
        elif self.distributed:
            segs = None
            if self.capture_collectives:
                try:
                    cap([mfcc, message, self.std.exchange, finish])
                    segs = []
                except Exception:                       # noqa: BLE001  (capture errors surface as RuntimeError subclasses)
                    sg.graphs.clear()
                    self.capture_collectives = False
                    torch.cuda.synchronize(self.device)
            if segs is None:
                segs = [[mfcc, message], [finish]]
                sg.exchange_after = 0
        else:
            segs = [[mfcc, lambda: self.std.fit_transform(flat, out=out)]]
        for seg in segs:
            cap(seg)
        return sg

    def run_corpus(self, batches, n_local: int, out_dtype=torch.float32):
        """BASELINE configs[3]: this rank's shard of a corpus arrives batch by batch (``batches`` yields
        ``(ClipBatch, z or None, snr_db or None)``); the features of all batches stay resident in one ``(n_local, D)``
        matrix (1 M C1 clips = 5.3 GB), the column statistics are all-reduced ONCE for the whole corpus
        (``n_total`` = sum of ``n_local`` over the ranks), then every row is standardised."""
        feats = torch.empty((n_local, self.rows, self.out_frames), dtype=torch.float32, device=self.device)
        done = 0
        it = iter(batches)
        cur = next(it, None)
        if cur is not None and cur[2] is not None:
            self.prefetch_power(cur[0])
        while cur is not None:
            nxt = next(it, None)                             # one batch of look-ahead: its power pass goes in front
            batch, z, snr_db = cur
            sigma = None
            ahead = nxt[0] if nxt is not None and nxt[2] is not None else None
            if ahead is not None and self._pow_stream is not None:
                self._pow_entry.record(torch.cuda.current_stream(self.device))
            if snr_db is not None and self.sigma_mode == "host":
                sigma, _ = self._sigma_for(batch, snr_db, ahead)
            self._group1(batch, z, snr_db, feats[done:done + batch.n_clips], sigma)
            if ahead is not None and self._pow_stream is not None:
                self.prefetch_power(ahead, after=self._pow_entry)   # beside this batch's MFCC launch
            done += batch.n_clips
            cur = nxt
        if done != n_local:
            raise ValueError(f"the batches held {done} clips, expected {n_local}")
        flat = feats.view(n_local, self.D)
        return self.std.fit_transform(flat, out_dtype=out_dtype)

    def run_host(self, audio_host: torch.Tensor, snr_db: Optional[float], seed: int, out_host: torch.Tensor,
                 first_index: int = 0, layout: Optional[ClipBatch] = None, noise_kind: str = "white",
                 next_audio_host: Optional[torch.Tensor] = None) -> torch.Tensor:
        """End to end from PINNED host memory: (B, L) int16/float32/float64 host tensor in - or, with ``layout`` (a
        ``ClipBatch`` whose offsets / lengths describe it), a packed 1-D host tensor of ragged clips -, standardised
        float32 (B, D) rows written to the pinned `out_host`.  The noise stream is generated on the device from `seed`
        (element index = first_index + position in this shard).

        Three streams (host->device copy, compute, device->host copy) and two sets of device buffers: the
        upload of call i+1 overlaps the kernels of call i and the download of call i-1.  The caller's current
        stream is made to wait for this call's download, so `torch.cuda.current_stream().synchronize()` (or
        an event recorded after the call) covers it.

        ``next_audio_host``: the host tensor of the NEXT call (same shape / dtype / layout; its samples must be in place now
        and stay untouched until that call).  Its upload is enqueued BEFORE this call waits for its own sigma (in host
        sigma mode the call blocks on the power of its batch, i.e. on its own upload), and its power pass behind this
        call's kernels - the copy engine then never idles between calls and the next call finds its power ready.  Without
        it every call pays upload -> power -> host chain serially before the next upload can start."""
        hs = self._host_slots(audio_host)
        hkey = (audio_host.data_ptr(), tuple(audio_host.shape), audio_host.dtype)
        k = self._host_pre.pop(hkey, None)                 # already uploaded by the previous call?
        if k is None:
            k = self._host_turn
            self._upload(hs[k], audio_host)
        self._host_pre = {key: v for key, v in self._host_pre.items() if v != k}   # a stale announcement loses its slot
        self._host_turn = k ^ 1
        sl = hs[k]
        cur = torch.cuda.current_stream(self.device)
        babble = noise_kind == "babble"
        nxt = None
        if next_audio_host is not None:
            if next_audio_host.shape != audio_host.shape or next_audio_host.dtype != audio_host.dtype:
                raise ValueError("next_audio_host must have the shape and dtype of audio_host")
            nxt = hs[k ^ 1]
            self._upload(nxt, next_audio_host)             # waits (on the device) for the kernels that last read that slot
            self._host_pre[(next_audio_host.data_ptr(), tuple(next_audio_host.shape), next_audio_host.dtype)] = k ^ 1
        self._s_comp.wait_event(sl["uploaded"])
        self._s_comp.wait_event(sl["downloaded"])          # this slot's output buffer has been read back
        with torch.cuda.stream(self._s_comp):
            batch = ClipBatch.from_matrix(sl["audio"]) if layout is None else layout.like(sl["audio"])
            z = None
            if snr_db is not None and (self.sigma_mode == "host" or babble) and self._pow_stream is not None \
                    and not (self._sig is not None and self._sig["pending"].get(self._bkey(batch) + (babble,))):
                self.prefetch_power(batch, babble)         # not announced: the power pass beside (not behind) the noise generator
            if snr_db is not None and noise_kind == "white":
                if sl["z"] is None:
                    sl["z"] = torch.empty(sl["audio"].numel(), dtype=torch.float64, device=self.device)
                z = sl["z"]
                randn(seed, first_index, z.numel(), device=self.device, out=z)
            out = self.run_device(batch, z, snr_db, noise_kind=noise_kind)
            sl["computed"].record()
            if nxt is not None and snr_db is not None:
                # the next call's power pass: behind this call's kernels, as soon as its samples have arrived
                self._s_comp.wait_event(nxt["uploaded"])
                self.prefetch_power(ClipBatch.from_matrix(nxt["audio"]) if layout is None else layout.like(nxt["audio"]), babble)
        self._s_d2h.wait_event(sl["computed"])
        with torch.cuda.stream(self._s_d2h):
            out_host.copy_(out, non_blocking=True)
            # `out` (and the feature buffer behind it) belong to the compute stream's allocator: keep them alive until
            # this copy has run, whatever graph mode the step used
            out.record_stream(self._s_d2h)
            sl["downloaded"].record()
        cur.wait_event(sl["downloaded"])
        return out_host

    def _upload(self, sl, audio_host: torch.Tensor) -> None:
        if self._sig is not None:                          # a power pass of the slot's OLD content must not be taken for the new one
            ptr = sl["audio"].data_ptr()
            for key in [key for key in self._sig["pending"] if key[0] == ptr]:
                del self._sig["pending"][key]
        self._s_h2d.wait_event(sl["computed"])            # the kernels of the call that last used this slot are done
        with torch.cuda.stream(self._s_h2d):
            sl["audio"].copy_(audio_host, non_blocking=True)
            sl["uploaded"].record()

    def _host_slots(self, audio_host: torch.Tensor):
        hs = getattr(self, "_hs", None)
        if hs is None or hs[0]["audio"].shape != audio_host.shape or hs[0]["audio"].dtype != audio_host.dtype:
            self._s_h2d, self._s_comp, self._s_d2h = (torch.cuda.Stream(self.device) for _ in range(3))
            hs = []
            for _ in range(2):
                hs.append({"audio": torch.empty(audio_host.shape, dtype=audio_host.dtype, device=self.device),
                           "z": None,          # float64 noise stream, allocated by the first white-noise call
                           "uploaded": torch.cuda.Event(), "computed": torch.cuda.Event(), "downloaded": torch.cuda.Event()})
            self._hs = hs
            self._host_turn = 0
            self._host_pre = {}                            # host tensor announced by `next_audio_host` -> slot holding it
        return hs
