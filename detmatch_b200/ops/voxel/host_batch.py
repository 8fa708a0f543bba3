"""Host-buffer entry point of batched hard voxelization: the reference's ``voxelization()`` accepts CPU
tensors (mmdet3d/ops/voxel/src/voxelization.h:66-68 dispatches on ``points.device()``) and the
detectors' ``voxelize()`` loop (openpcdet.py:59-76, voxelnet.py:50-67) turns a list of frames into the
concatenated ``(voxels, num_points, coors_batch)``.  ``voxelize_batch_host`` does that for a list of
HOST frames with the results in (pinned) HOST memory: every frame is uploaded, voxelized on the GPU --
there is no CPU code path -- and the concatenated rows are read back.

The batch is cut into chunks of ``chunk`` frames (the first one shorter) on three streams: the upload of chunk i + 1, the
kernels of chunk i and the read-back of chunk i - 1 overlap.  A chunk runs the packed C-ABI call
(``pcfe_hard_voxelize_packed_batch_f32``: concatenated rows written by the expansion kernel itself, no
``torch.cat`` staging), so one chunk leaves the device with three copies that land directly at their
final offset of the batch's output tensors.  The size of a chunk's read-back is its voxel count:
one 4 * chunk byte copy and one event wait per chunk, issued while later chunks are still computing.
"""
import ctypes

import torch

from ... import _cabi
from ..._torch_glue import ptr, require_cuda
from .voxelize import voxelize_batch


class HostVoxelizePipeline:
    """Pre-allocated pipeline for a fixed batch geometry (frame sizes, C, grid, caps): device input
    and output staging, pinned host outputs, streams.  ``run(host_points)`` returns host views
    ``(voxels (sum M, P, C), num_points (sum M,), coors_batch (sum M, 4))`` that stay valid until the
    next ``run``."""

    def __init__(self, sizes, num_features, voxel_size, coors_range, max_points, max_voxels, device=None, chunk=8):
        require_cuda()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        dev = self.device
        self.sizes = [int(n) for n in sizes]
        self.c, self.p, self.v = int(num_features), int(max_points), int(max_voxels)
        self.voxel_size, self.coors_range = list(voxel_size), list(coors_range)
        self.vs, self.rg = _cabi.f3(voxel_size), _cabi.f6(coors_range)
        F = len(self.sizes)
        # a short first chunk: nothing overlaps its upload, so the pipeline fills in a quarter of the time
        # (64 C4 frames, chunk 8: chunks of 2, 6, 8, 8, ...)
        first = max(1, chunk // 4) if F > chunk else chunk
        cuts = sorted({0, min(first, F), *range(chunk, F, chunk), F})
        self.chunks = [list(range(a, b)) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        self.packed = self.p == 5 and self.c in (4, 5)
        self.dpts = [torch.empty((n, self.c), dtype=torch.float32, device=dev) for n in self.sizes]
        caps = [sum(min(self.sizes[k], self.v) for k in ch) for ch in self.chunks]
        capc = max(caps + [1])
        self.stage = [(torch.empty((capc, self.p, self.c), dtype=torch.float32, device=dev),
                       torch.empty((capc,), dtype=torch.int32, device=dev),
                       torch.empty((capc, 4), dtype=torch.int32, device=dev),
                       torch.empty((chunk,), dtype=torch.int32, device=dev)) for _ in range(2)]
        cap_total = max(sum(caps), 1)
        self.out_vox = torch.empty((cap_total, self.p, self.c), dtype=torch.float32).pin_memory()
        self.out_num = torch.empty((cap_total,), dtype=torch.int32).pin_memory()
        self.out_coors = torch.empty((cap_total, 4), dtype=torch.int32).pin_memory()
        self.cnt_host = [torch.empty((chunk,), dtype=torch.int32).pin_memory() for _ in self.chunks]
        self.s_in, self.s_comp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        L = _cabi.lib()
        need = L.pcfe_hard_voxelize_workspace_bytes(max(self.sizes + [0]), chunk, 0, self.vs, self.rg, self.p, self.v)
        self.ws = torch.empty(max(int(need), 256), dtype=torch.uint8, device=dev)
        self.h2d_bytes = sum(self.sizes) * self.c * 4
        self.d2h_bytes = 0
        self.counts = []

    def _compute_chunk(self, i, ch, buf):
        """Enqueues the voxelization of chunk i on the current (compute) stream; rows land in `buf`."""
        sv, sn, sc, cnt = buf
        L = _cabi.lib()
        nf = len(ch)
        if self.packed:
            pp = (ctypes.c_void_p * nf)(*[self.dpts[k].data_ptr() for k in ch])
            nn = (ctypes.c_int64 * nf)(*[self.sizes[k] for k in ch])
            rc = L.pcfe_hard_voxelize_packed_batch_f32(pp, nn, nf, self.c, self.vs, self.rg, None, self.p, self.v, 0,
                                                       ptr(sv), ptr(sc), ptr(sn), sv.size(0), ptr(cnt), ptr(self.ws),
                                                       self.ws.numel(), self.device.index,
                                                       torch.cuda.current_stream(self.device).cuda_stream)
            if rc == 0:
                if ch[0]:
                    sc[:, 0] += ch[0]  # batch index of the chunk's first frame (rows past the end are never read)
                return None
            if rc != _cabi.ERR_SHAPE:
                _cabi.check(rc, "pcfe_hard_voxelize_packed_batch_f32")
        # shapes without packed output: per-frame outputs concatenated on the device (one host read inside)
        v, n, cb = voxelize_batch([self.dpts[k] for k in ch], self.voxel_size, self.coors_range, self.p, self.v, sync=True)
        tot = v.size(0)
        sv[:tot].copy_(v)
        sn[:tot].copy_(n)
        sc[:tot].copy_(cb)
        if ch[0]:
            sc[:tot, 0] += ch[0]
        per = torch.bincount(cb[:, 0].long(), minlength=nf).int() if tot else torch.zeros((nf,), dtype=torch.int32, device=self.device)
        cnt[:nf].copy_(per)
        return None

    def run(self, host_points):
        assert len(host_points) == len(self.sizes)
        dev = self.device
        ev_comp, ev_out = [], [None, None]
        start = torch.cuda.Event()
        start.record(torch.cuda.current_stream(dev))
        for s in (self.s_in, self.s_comp, self.s_out):
            s.wait_event(start)
        for i, ch in enumerate(self.chunks):
            with torch.cuda.stream(self.s_in):
                for k in ch:
                    h = host_points[k]
                    assert h.dtype == torch.float32 and tuple(h.shape) == (self.sizes[k], self.c) and not h.is_cuda
                    self.dpts[k].copy_(h, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(ev_in)
                if ev_out[i & 1] is not None:
                    self.s_comp.wait_event(ev_out[i & 1])  # the staging buffer's previous read-back is done
                self._compute_chunk(i, ch, self.stage[i & 1])
                self.cnt_host[i].copy_(self.stage[i & 1][3], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.s_comp)
                ev_comp.append(ev)
            # read-backs are enqueued one chunk behind, so that the compute stream never waits for the host
            if i >= 1:
                ev_out[(i - 1) & 1] = self._read_back(i - 1, ev_comp[i - 1])
        if self.chunks:
            self._read_back(len(self.chunks) - 1, ev_comp[-1])
        self.s_out.synchronize()
        total = self._row0
        return self.out_vox[:total], self.out_num[:total], self.out_coors[:total]

    def _read_back(self, i, ev):
        if i == 0:
            self._row0 = 0
            self.d2h_bytes = 0
            self.counts = []
        ev.synchronize()  # the chunk's voxel counts size its read-back
        nf = len(self.chunks[i])
        counts = self.cnt_host[i][:nf].tolist()
        tot = sum(counts)
        sv, sn, sc, _ = self.stage[i & 1]
        r0 = self._row0
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(ev)
            self.out_vox[r0:r0 + tot].copy_(sv[:tot], non_blocking=True)
            self.out_num[r0:r0 + tot].copy_(sn[:tot], non_blocking=True)
            self.out_coors[r0:r0 + tot].copy_(sc[:tot], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.s_out)
        self._row0 = r0 + tot
        self.counts.extend(counts)
        self.d2h_bytes += nf * 4 + tot * (self.p * self.c * 4 + 4 + 16)
        return done


def voxelize_batch_host(points, voxel_size, coors_range, max_points, max_voxels, device=None, chunk=8):
    """List of HOST (N_i, C) float32 frames -> ``(voxels, num_points, coors_batch)`` in pinned host
    memory, exactly the concatenated tensors of the reference's ``voxelize()`` loop.  One-shot form of
    ``HostVoxelizePipeline`` (which a loop over same-shaped batches should keep and reuse)."""
    assert len(points) > 0
    pts = [p.contiguous() for p in points]
    pipe = HostVoxelizePipeline([p.size(0) for p in pts], pts[0].size(1), voxel_size, coors_range, max_points, max_voxels,
                                device=device, chunk=chunk)
    v, n, c = pipe.run(pts)
    return v.clone(), n.clone(), c.clone()
