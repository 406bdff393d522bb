"""End-to-end latent-walk edit: host latents in, host uint8 images out.

This is the call a user of the accelerated path makes for the forward-only edit workload
(BASELINE.json configs[1]): ``z -> w = style(z) -> w' = walk(w, alpha) -> G(w') -> uint8 panels``,
i.e. the body of the reference's ``apply_alpha`` + the host-side ``clip((x+1)/2*255)`` of
``vis_multi_image_batch_alphas`` (transform_base.py:554-603, 625-626) with the attribute regressor
factored out (``alpha`` is the walk step epsilon).  Host buffers are
pinned once; every call copies z / alpha to the device on the current stream and the uint8 result
back on a dedicated copy stream (double-buffered), so the 3 MB/image device->host transfer of step i
overlaps the kernels of step i+1.  The per-layer noise of call i+1 (17 Philox launches, 358 MB at 1024 px x 32) is drawn
on a side stream while call i's convolutions run: same generator, same draw order and shapes as
``NoiseInjection`` (networks.py:281-286), so a seeded run still consumes the reference's random stream.
"""
from __future__ import annotations

import numpy as np
import torch


class EditPipeline:
    def __init__(self, generator, walk, batch: int, n_attr: int = 1, device=None):
        self.gen, self.walk, self.batch = generator, walk, batch
        self.device = torch.device(device) if device is not None else generator.input.input.device
        if self.device.type != "cuda":
            raise RuntimeError("EditPipeline needs a CUDA device (there is no CPU fallback)")
        dim, size = generator.style_dim, generator.size
        # inputs are double-buffered like the outputs: with sync=False the host may fill buffer k^1 while the
        # host->device copy out of buffer k is still queued behind the previous call's kernels
        self.z_hosts = [torch.empty(batch, dim, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.alpha_hosts = [torch.empty(batch, n_attr, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.h2d_done = [None, None]       # event: host->device copies out of pinned input buffer k have finished
        self.out_hosts = [torch.empty(batch, size, size, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.out_devs = [torch.empty(batch, size, size, 3, dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(self.device)
        self.noise_stream = torch.cuda.Stream(self.device)
        self._next_noise = None            # (tensors, ready event, set index) drawn ahead for the next call
        self._noise_in_use = None
        self._noise_pool, self._noise_pool_batch, self._noise_free, self._noise_j = None, None, None, 0
        self.copy_done = [None, None]      # event: device->host copy out of buffer k has finished
        self._k = 0
        self.z_devs = [torch.empty(batch, dim, dtype=torch.float32, device=self.device) for _ in range(2)]
        self.alpha_devs = [torch.empty(batch, n_attr, dtype=torch.float32, device=self.device) for _ in range(2)]
        generator.set_native(max_batch=batch)

    @property
    def h2d_bytes(self) -> int:
        return self.z_hosts[0].numel() * 4 + self.alpha_hosts[0].numel() * 4

    @property
    def d2h_bytes(self) -> int:
        return self.out_hosts[0].numel()

    _NOISE_SETS = 3   # one in use by the kernels in flight, one drawn ahead, one spare for a host that runs ahead

    def _noise_set(self, batch, j):
        """Preallocated per-layer noise tensors of set ``j`` (no allocation in steady state: a host enqueueing many
        steps ahead of the GPU would otherwise make the caching allocator cudaMalloc a fresh 11 MB/image set per step)."""
        if self._noise_pool is None or self._noise_pool_batch != batch:
            self._noise_pool = [[torch.empty(batch, 1, 2 ** ((i + 5) // 2), 2 ** ((i + 5) // 2), device=self.device, dtype=torch.float32)
                                 for i in range(self.gen.num_layers)] for _ in range(self._NOISE_SETS)]
            self._noise_free = [None] * self._NOISE_SETS   # event: the kernels that read set j have finished
            self._noise_pool_batch, self._noise_j = batch, 0
        return self._noise_pool[j]

    def _draw_noise_ahead(self, batch):
        self._noise_set(batch, 0)
        j = self._noise_j
        self._noise_j = (j + 1) % self._NOISE_SETS
        nz = self._noise_set(batch, j)
        with torch.cuda.stream(self.noise_stream):
            if self._noise_free[j] is not None:
                self.noise_stream.wait_event(self._noise_free[j])
            for t in nz:
                t.normal_()
            ev = torch.cuda.Event()
            ev.record(self.noise_stream)
        return nz, ev, j

    def _take_noise(self, batch):
        """Fresh per-layer noise for this call (drawn ahead on the side stream when a previous call left one); the
        caller starts the draw for the next call behind the kernels it enqueues."""
        main = torch.cuda.current_stream(self.device)
        if self._next_noise is None or self._next_noise[0][0].shape[0] != batch:
            self.noise_stream.wait_stream(main)    # a first draw must not overtake earlier work on the main stream
            self._next_noise = self._draw_noise_ahead(batch)
        nz, ev, j = self._next_noise
        main.wait_event(ev)
        self._noise_in_use = j
        self._next_noise = None
        return nz

    def _release_noise(self):
        """Marks the set taken by the current call as free once the kernels enqueued so far have run."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._noise_free[self._noise_in_use] = ev

    @torch.no_grad()
    def edit_device(self, z_dev, alpha_dev, layers=None, noise=None, want_uint8=False, out_uint8=None):
        """Device-resident part: mapping -> walk -> synthesis.  Returns the fp32 image (reference
        return value) or the uint8 NHWC image."""
        w = self.gen.style(z_dev)
        ws = self.walk([w] * self.gen.n_latent, alpha_dev, layers=layers)
        latent = torch.stack(ws, 1) if not _is_block(ws) else ws[0]._base
        drew = noise is None
        if drew:
            noise = self._take_noise(latent.shape[0])
        out = self.gen.synthesize(latent, noise=noise, want_uint8=want_uint8, want_float=not want_uint8, out_uint8=out_uint8)
        if drew:   # the next call's noise is generated while the kernels just enqueued run
            self._release_noise()
            self._next_noise = self._draw_noise_ahead(latent.shape[0])
        return out

    @torch.no_grad()
    def edit(self, z, alpha, layers=None, noise=None, sync=True) -> np.ndarray:
        """``z``: [B, dim] host array / tensor (float64 from the numpy sampler is fine, it is cast to
        float32 exactly like ``torch.Tensor(z)`` in train.py:56); ``alpha``: [B, A] host."""
        main = torch.cuda.current_stream(self.device)
        k = self._k
        self._k ^= 1
        if self.h2d_done[k] is not None:
            self.h2d_done[k].synchronize()               # the copy engine has read pinned input buffer k (two calls ago)
        self.z_hosts[k].copy_(torch.as_tensor(z).to(torch.float32))
        self.alpha_hosts[k].copy_(torch.as_tensor(alpha).to(torch.float32).reshape(self.batch, -1))
        if self.copy_done[k] is not None:
            main.wait_event(self.copy_done[k])           # output buffer k is free again (copy of two calls ago)
        self.z_devs[k].copy_(self.z_hosts[k], non_blocking=True)
        self.alpha_devs[k].copy_(self.alpha_hosts[k], non_blocking=True)
        self.h2d_done[k] = torch.cuda.Event()
        self.h2d_done[k].record(main)
        self.edit_device(self.z_devs[k], self.alpha_devs[k], layers=layers, noise=noise, want_uint8=True, out_uint8=self.out_devs[k])
        computed = torch.cuda.Event()
        computed.record(main)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(computed)
            self.out_hosts[k].copy_(self.out_devs[k], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self.copy_done[k] = done
        if sync:
            done.synchronize()
        return self.out_hosts[k].numpy()

    def join(self):
        """Makes the current stream wait for every outstanding device->host copy (no host blocking)."""
        main = torch.cuda.current_stream(self.device)
        for ev in self.copy_done:
            if ev is not None:
                main.wait_event(ev)

    def wait(self):
        """Blocks until every outstanding device->host copy has landed."""
        for ev in self.copy_done:
            if ev is not None:
                ev.synchronize()


def _is_block(ws) -> bool:
    """True when the walk returned the unbound views of one contiguous [B, L, D] tensor, in order."""
    base = getattr(ws[0], "_base", None)
    if base is None or base.ndim != 3 or base.shape[1] != len(ws):
        return False
    return all(getattr(w, "_base", None) is base and w.data_ptr() == base[:, i].data_ptr() for i, w in enumerate(ws))
