"""Device engine: thin Python layer over the C ABI (``include/cmb200.h``).

PyTorch is used only for device memory, streams and (in ``parallel.py``)
``torch.distributed``; every computation is a hand-written sm_100a kernel behind
``libcmb200.so``.  There is no CPU fallback: constructing an engine without a
CUDA device raises RuntimeError.
"""
import ctypes
import time

import numpy as np
import torch

from . import _lib


class EngineError(RuntimeError):
    pass


def _ptr(t):
    """Raw address of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)


def _stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class HashTable(object):
    """Hashed count table of the global-memory cell path (``cmb_set_table_mode``): 2^log2 uint64
    slots, slot = (flat index + 1) << 24 | count, zero = empty.  ``used`` tracks the number of
    distinct keys (the engine keeps the load below 0.85 by rehashing into a larger table: small
    tables stay L2-resident, which is the point of hashing);
    ``merged`` holds the (index, count) arrays after a multi-rank reduce."""

    def __init__(self, log2, device):
        self.log2 = int(log2)
        self.slots = torch.zeros(1 << self.log2, dtype=torch.int64, device=device)
        self.used = 0
        self.merged = None

    def numel(self):
        return self.slots.numel()

    def zero_(self):
        self.slots.zero_()
        self.used = 0
        self.merged = None
        return self


def _table_ptr(t):
    return _ptr(t.slots) if isinstance(t, HashTable) else _ptr(t)


def _table_log2(t):
    return t.log2 if isinstance(t, HashTable) else 0


class ContactEngine(object):
    """One engine per (process, device).  Not thread safe."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise EngineError("contact_map_b200 needs a CUDA device (sm_100a); there is no CPU "
                              "fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", int(device) if not isinstance(device, torch.device)
                                   else device.index or 0)
        self._L = _lib.lib()
        handle = ctypes.c_void_p()
        rc = self._L.cmb_ctx_create(self.device.index, ctypes.byref(handle))
        if rc != 0:
            raise EngineError("cmb_ctx_create failed (%d): %s"
                              % (rc, self._L.cmb_last_error(None).decode()))
        self._ctx = handle
        self.n_used = 0
        self.n_res_c = 0
        self.residue_map = None
        self._system_key = None
        self._keepalive = []

    # ------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._L.cmb_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:     # interpreter shutdown
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise EngineError("%s failed (%d): %s"
                              % (what, rc, self._L.cmb_last_error(self._ctx).decode()))

    # ------------------------------------------------------------------
    def set_option(self, name, value):
        self._check(self._L.cmb_set_option(self._ctx, name.encode(), int(value)), "cmb_set_option")
        if name == "force_path":
            self._forced_path = int(value)
        if name in ("force_path", "max_cells"):
            self._system_key = None           # these take effect at the next cmb_set_system

    def get_stat(self, name):
        out = ctypes.c_int64(0)
        self._check(self._L.cmb_get_stat(self._ctx, name.encode(), ctypes.byref(out)), "cmb_get_stat")
        return int(out.value)

    def set_system(self, atom_res, atom_chain, role, cutoff, n_ignored):
        atom_res = np.ascontiguousarray(atom_res, dtype=np.int32)
        atom_chain = np.ascontiguousarray(atom_chain, dtype=np.int32)
        role = np.ascontiguousarray(role, dtype=np.uint8)
        n = atom_res.shape[0]
        if atom_chain.shape[0] != n or role.shape[0] != n:
            raise ValueError("system arrays must have the same length")
        # the same system again (one analysis object per trajectory chunk): nothing to upload
        key = (n, float(cutoff), int(n_ignored), atom_res.tobytes(), atom_chain.tobytes(), role.tobytes(),
               self.get_stat("max_cells"), int(getattr(self, "_forced_path", -1)))
        if key == getattr(self, "_system_key", None):
            return
        self._system_key = None
        self._check(self._L.cmb_set_system(self._ctx, n, _ptr(atom_res), _ptr(atom_chain),
                                           _ptr(role), float(cutoff), int(n_ignored)),
                    "cmb_set_system")
        self.n_used = n
        self.n_res_c = int(self._L.cmb_n_residues(self._ctx))
        rmap = np.empty(self.n_res_c, dtype=np.int32)
        self._check(self._L.cmb_residue_map(self._ctx, _ptr(rmap)), "cmb_residue_map")
        self.residue_map = rmap
        self.path_kind = int(self._L.cmb_path_kind(self._ctx))
        self._system_key = key

    # dense tables above this size are replaced by hashed ones (global-memory cell path only).
    # Measured on the 100k-atom system (40 GB dense table): dense 6.3k frames/s, hashed 5.4k --
    # the hashed insert is a dependent L2 round trip where the dense path issues a RED -- so dense
    # stays the default while it fits comfortably in the 180 GB of a B200.
    DENSE_LIMIT_BYTES = 1 << 36

    def new_tables(self, hashed=None):
        """Zeroed device count tables: dense int32 views of the uint32 tables of the ABI, or
        (``hashed=True``; default for systems whose dense atom table would exceed 64 GiB)
        :class:`HashTable` objects."""
        n, rc = int(self.n_used), int(self.n_res_c)
        if hashed is None:
            hashed = self.path_kind == 1 and 4 * n * n > self.DENSE_LIMIT_BYTES
        if hashed:
            if self.path_kind != 1:
                raise EngineError("hashed count tables are only used by the global-memory cell path")
            atom = HashTable(max(10, int(np.ceil(np.log2(32.0 * n)))), self.device)
            if 4 * rc * rc > (1 << 28):
                res = HashTable(max(10, int(np.ceil(np.log2(48.0 * rc)))), self.device)
            else:
                res = torch.zeros(rc * rc, dtype=torch.int32, device=self.device)
            return atom, res
        # one allocation, two views (16-byte aligned): the multi-rank reduce is then a single
        # all-reduce of the allocation, with no packing copies (parallel.reduce_tables)
        n_atom = (n * n + 3) // 4 * 4
        flat = torch.zeros(n_atom + rc * rc, dtype=torch.int32, device=self.device)
        return flat[:n * n], flat[n_atom:n_atom + rc * rc]

    # ---- hashed tables: mode switch, frame schedule, growth ---------------------------
    def _set_table_mode(self, atom_table, res_table):
        mode = (_table_log2(atom_table), _table_log2(res_table))
        if mode != getattr(self, "_table_mode", (0, 0)):
            self._check(self._L.cmb_set_table_mode(self._ctx, mode[0], mode[1]), "cmb_set_table_mode")
            self._table_mode = mode
        if mode != (0, 0) and getattr(self, "_ovf", None) is None:
            self._ovf = torch.zeros(self.OVERFLOW_ENTRIES, dtype=torch.int64, device=self.device)
            self._check(self._L.cmb_hash_overflow_buffer(self._ctx, _ptr(self._ovf), self.OVERFLOW_ENTRIES),
                        "cmb_hash_overflow_buffer")
        return mode != (0, 0)

    @staticmethod
    def _hashed_schedule(n_frames, cap=256):
        """Frame blocks 1, 1, 2, 4, ... (at most ``cap``): the tables are checked, and grown when
        more than half full, after each block; the pair universe of a trajectory saturates after a
        few frames, so a block never comes close to filling a table that was at most half full."""
        out, start, size = [], 0, 1
        while start < n_frames:
            stop = min(n_frames, start + size)
            out.append((start, stop))
            start = stop
            size = min(cap, max(1, 2 * size if len(out) > 1 else 1))
        return out

    def _hash_presize(self, xyz1, box1, atom_table, res_table):
        """Fresh hashed tables are sized from the exact pair counts of the first frame (a counting
        pass of the key kernel that writes nothing), so that the first block cannot overflow."""
        if not (isinstance(atom_table, HashTable) and atom_table.used == 0):
            return
        self._set_table_mode(None, None)
        counts = self._frame_contacts_once(xyz1, box1, 0, None, 0, None, 0, None, None,
                                           count_only=(True, False))
        # (a frame has at most as many residue pairs as atom pairs, usually far fewer: the residue
        # table keeps its default size, 64 slots per residue, unless even that bound is smaller)
        for t, first in ((atom_table, int(counts[0])),):
            if isinstance(t, HashTable) and t.used == 0 and 2 * first > t.numel():
                log2 = t.log2
                while 2 * first > (1 << log2):
                    log2 += 1
                t.slots, t.log2 = torch.zeros(1 << log2, dtype=torch.int64, device=self.device), log2
        self._set_table_mode(atom_table, res_table)

    def _hash_presize_host(self, xyz1_np, box_np, atom_table, res_table):
        xyz1 = torch.as_tensor(xyz1_np, device=self.device)
        box1 = None if box_np is None else torch.as_tensor(box_np[:1], device=self.device)
        self._hash_presize(xyz1, box1, atom_table, res_table)

    OVERFLOW_ENTRIES = 1 << 22          # parked hits per block (32 MiB list)

    def _grow(self, t, min_slots):
        log2 = t.log2
        while (1 << log2) < min_slots:
            log2 += 1
        if log2 == t.log2:
            return
        bigger = torch.zeros(1 << log2, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._L.cmb_hash_rehash(self._ctx, _ptr(t.slots), t.log2, _ptr(bigger), log2,
                                                _stream_ptr(self.device)), "cmb_hash_rehash")
        torch.cuda.current_stream(self.device).synchronize()
        t.slots, t.log2 = bigger, log2

    def _hash_maintain(self, tables):
        """After a block of frames: bookkeeping of the distinct keys, growth (load <= 0.7), and the
        hits that found their table full (parked in the overflow list by the kernel): the table
        is grown and they are added, per-frame duplicates of a residue pair removed first."""
        a, r, o = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int64(0)
        self._check(self._L.cmb_hash_stats(self._ctx, ctypes.byref(a), ctypes.byref(r), ctypes.byref(o)),
                    "cmb_hash_stats")
        parked = int(o.value)
        if parked > self.OVERFLOW_ENTRIES:
            raise EngineError("%d hits overflowed the hashed count tables and their overflow list "
                              "(hits were lost); use larger tables" % parked)
        adds = [None, None]
        if parked:
            entries = self._ovf[:parked]
            is_res = entries < 0                                  # bit 63
            atoms = entries[~is_res]
            if atoms.numel():
                adds[0] = torch.unique(atoms, return_counts=True)
            res = entries[is_res]
            if res.numel():
                per_frame = torch.unique(res)                     # one entry per (pair, frame)
                keys = (per_frame & 0x7fffffffffffffff) >> 24
                adds[1] = torch.unique(keys, return_counts=True)
        for which, (t, new) in enumerate(zip(tables, (a.value, r.value))):
            if not isinstance(t, HashTable):
                continue
            t.used += int(new)
            t.merged = None
            extra = 0 if adds[which] is None else int(adds[which][0].numel())
            # load > 0.85: grow to load <= 0.6.  Bucketed probing tolerates high loads, and a SMALL
            # table is the point: the hot set (tables + stamps) has to stay inside the 126 MB L2
            if 100 * (t.used + extra) > 85 * t.numel():
                self._grow(t, int((t.used + extra) / 0.6) + 1)
            if extra:
                keys, counts = adds[which]
                with torch.cuda.device(self.device):
                    self._check(self._L.cmb_hash_add(self._ctx, _ptr(t.slots), t.log2, _ptr(keys.contiguous()),
                                                     _ptr(counts.contiguous()), extra, which,
                                                     _stream_ptr(self.device)), "cmb_hash_add")
        if parked:
            self._check(self._L.cmb_hash_stats(self._ctx, ctypes.byref(a), ctypes.byref(r), ctypes.byref(o)),
                        "cmb_hash_stats")
            for t, new in zip(tables, (a.value, r.value)):
                if isinstance(t, HashTable):
                    t.used += int(new)

    # ------------------------------------------------------------------
    def _dev_inputs(self, xyz, box):
        if not (isinstance(xyz, torch.Tensor) and xyz.is_cuda):
            raise TypeError("device path needs CUDA tensors")
        if xyz.dtype != torch.float32 or not xyz.is_contiguous():
            raise TypeError("xyz must be a contiguous float32 tensor")
        if xyz.dim() != 3 or xyz.shape[1] != self.n_used or xyz.shape[2] != 3:
            raise ValueError("xyz must have shape [n_frames, %d, 3]" % self.n_used)
        if box is not None:
            if box.dtype != torch.float32 or not box.is_contiguous() or not box.is_cuda:
                raise TypeError("box must be a contiguous float32 CUDA tensor")
            if box.shape[0] != xyz.shape[0] or box.numel() != xyz.shape[0] * 9:
                raise ValueError("box must have shape [n_frames, 3, 3]")
        return xyz, box

    def accumulate(self, xyz, box, atom_table, res_table):
        """Add the frames of a DEVICE trajectory into the tables (asynchronous)."""
        xyz, box = self._dev_inputs(xyz, box)
        if self._set_table_mode(atom_table, res_table):
            # hashed tables: block schedule with growth checks in between (synchronous)
            self._hash_presize(xyz[:1], None if box is None else box[:1], atom_table, res_table)
            for a, b in self._hashed_schedule(xyz.shape[0]):
                with torch.cuda.device(self.device):
                    self._set_table_mode(atom_table, res_table)      # growth changes the sizes
                    self._check(self._L.cmb_accumulate(self._ctx, _ptr(xyz[a:b]),
                                                       _ptr(None if box is None else box[a:b]), b - a,
                                                       _table_ptr(atom_table), _table_ptr(res_table),
                                                       _stream_ptr(self.device)), "cmb_accumulate")
                self._hash_maintain((atom_table, res_table))
            return
        with torch.cuda.device(self.device):
            self._check(self._L.cmb_accumulate(self._ctx, _ptr(xyz), _ptr(box), xyz.shape[0],
                                               _ptr(atom_table), _ptr(res_table),
                                               _stream_ptr(self.device)), "cmb_accumulate")

    def accumulate_host(self, xyz, box, atom_table, res_table, chunk_frames=0):
        """Add the frames of a HOST trajectory (numpy or CPU tensor; pinned is fastest)."""
        xyz_np = xyz.numpy() if isinstance(xyz, torch.Tensor) else xyz
        box_np = box.numpy() if isinstance(box, torch.Tensor) else box
        if xyz_np.dtype != np.float32 or not xyz_np.flags["C_CONTIGUOUS"]:
            raise TypeError("xyz must be C-contiguous float32")
        if xyz_np.ndim != 3 or xyz_np.shape[1] != self.n_used or xyz_np.shape[2] != 3:
            raise ValueError("xyz must have shape [n_frames, %d, 3]" % self.n_used)
        if box_np is not None:
            if box_np.dtype != np.float32 or not box_np.flags["C_CONTIGUOUS"]:
                raise TypeError("box must be C-contiguous float32")
            if box_np.size != xyz_np.shape[0] * 9:
                raise ValueError("box must have shape [n_frames, 3, 3]")
        if self._set_table_mode(atom_table, res_table):
            torch.cuda.current_stream(self.device).synchronize()   # tables may still be zero-filling
            self._hash_presize_host(xyz_np[:1], box_np, atom_table, res_table)
            for a, b in self._hashed_schedule(xyz_np.shape[0]):
                self._set_table_mode(atom_table, res_table)
                self._check(self._L.cmb_accumulate_host(self._ctx, _ptr(xyz_np[a:b]),
                                                        _ptr(None if box_np is None else box_np[a:b]),
                                                        b - a, int(chunk_frames), _table_ptr(atom_table),
                                                        _table_ptr(res_table)), "cmb_accumulate_host")
                self._hash_maintain((atom_table, res_table))
            return
        # (the tables may still be zero-filling on torch's stream: the kernels wait for that, the
        # first host-to-device copies do not)
        self._check(self._L.cmb_wait_for_stream(self._ctx, _stream_ptr(self.device)), "cmb_wait_for_stream")
        self._check(self._L.cmb_accumulate_host(self._ctx, _ptr(xyz_np), _ptr(box_np),
                                                xyz_np.shape[0], int(chunk_frames),
                                                _ptr(atom_table), _ptr(res_table)),
                    "cmb_accumulate_host")

    def accumulate_host_gather(self, xyz_full, used, box, atom_table, res_table, chunk_frames=0,
                               n_threads=0):
        """Like :meth:`accumulate_host` for a HOST trajectory that still holds ALL atoms: the used
        atoms (ascending indices) are gathered by host threads into pinned staging while earlier
        chunks are copied and processed (fused ``_atom_slice``)."""
        xyz_np = xyz_full.numpy() if isinstance(xyz_full, torch.Tensor) else xyz_full
        box_np = box.numpy() if isinstance(box, torch.Tensor) else box
        if xyz_np.dtype != np.float32 or not xyz_np.flags["C_CONTIGUOUS"] or xyz_np.ndim != 3:
            raise TypeError("xyz must be C-contiguous float32 [n_frames, n_total, 3]")
        used = np.ascontiguousarray(used, dtype=np.int32)
        if used.shape[0] != self.n_used:
            raise ValueError("len(used) must equal the number of used atoms of the system")
        if box_np is not None and (box_np.dtype != np.float32 or not box_np.flags["C_CONTIGUOUS"]
                                   or box_np.size != xyz_np.shape[0] * 9):
            raise TypeError("box must be C-contiguous float32 [n_frames, 3, 3]")
        if self._set_table_mode(atom_table, res_table):
            torch.cuda.current_stream(self.device).synchronize()
            self._hash_presize_host(np.ascontiguousarray(xyz_np[:1][:, used]), box_np, atom_table, res_table)
            for a, b in self._hashed_schedule(xyz_np.shape[0]):
                self._set_table_mode(atom_table, res_table)
                self._check(self._L.cmb_accumulate_host_gather(
                    self._ctx, _ptr(xyz_np[a:b]), xyz_np.shape[1], _ptr(used),
                    _ptr(None if box_np is None else box_np[a:b]), b - a, int(chunk_frames), int(n_threads),
                    _table_ptr(atom_table), _table_ptr(res_table)), "cmb_accumulate_host_gather")
                self._hash_maintain((atom_table, res_table))
            return
        self._check(self._L.cmb_wait_for_stream(self._ctx, _stream_ptr(self.device)), "cmb_wait_for_stream")
        self._check(self._L.cmb_accumulate_host_gather(self._ctx, _ptr(xyz_np), xyz_np.shape[1], _ptr(used),
                                                       _ptr(box_np), xyz_np.shape[0], int(chunk_frames),
                                                       int(n_threads), _ptr(atom_table), _ptr(res_table)),
                    "cmb_accumulate_host_gather")

    def export_nonzero(self, table):
        """(flat index int64[], count uint32[]) of the non-zero entries, sorted by index."""
        return self.export_many([table])[0]

    def export_many(self, tables):
        """Sparse export of several count tables with ONE host synchronisation.

        Every table is scanned by ``cmb_export_sorted`` into a capacity-sized device buffer of
        packed (flat index << 24 | count) words (hashed tables: radix-sorted on the device), split
        into index and count arrays on the device (``cmb_unpack_sorted``, which reads the entry
        count there) and copied to pinned host memory asynchronously; the stream is synchronised
        once.  Capacities come from the previous export of a table of the same size (first use:
        ``count_nonzero``); an overflow triggers one exact retry; counts that do not fit 24 bits
        fall back to the unpacked export.
        Returns [(flat index int64[], count uint32[]), ...], each sorted by index.
        """
        fused = self._export_span(tables)
        if fused is not None:
            return fused
        hints = self.__dict__.setdefault("_export_hint", {})
        stream = torch.cuda.current_stream(self.device)
        jobs = []
        with torch.cuda.device(self.device):
            for t in tables:
                if isinstance(t, HashTable) and t.merged is not None:
                    jobs.append(None)
                    continue
                numel = t.numel()
                cap = t.used if isinstance(t, HashTable) else hints.get(numel)
                if cap is None:
                    cap = int(torch.count_nonzero(t).item())
                cap = max(16, min(numel, cap + cap // 4 + 1024))
                jobs.append(self._export_launch(t, cap, len(jobs)))
            stream.synchronize()
            out = []
            for slot, (t, job) in enumerate(zip(tables, jobs)):
                if job is None:
                    out.append(t.merged)
                    continue
                n, wide = int(job["h_n"][0]), int(job["h_n"][1])
                if wide:
                    out.append(self._export_unpacked(t))
                    continue
                if n > job["use"]:            # capacity guess too small: exact retry
                    job = self._export_launch(t, n, slot)
                    stream.synchronize()
                    n = int(job["h_n"][0])
                if not isinstance(t, HashTable):
                    hints[t.numel()] = n
                # (the pinned staging buffers are recycled by the next export: hand out copies)
                out.append((np.array(job["h_idx"][:n].numpy()), np.array(job["h_cnt"][:n].numpy()).view(np.uint32)))
        return out

    def export_pair_views(self, atom_table, res_table):
        """Sparse export of the two dense tables of :meth:`new_tables` as (row, column, count)
        triples split on the DEVICE (``cmb_unpack_pairs``): one scan of the allocation, one host
        synchronisation, no host-side divmod over millions of entries.

        Returns ``((lo, hi, cnt), (rlo, rhi, rcnt))`` -- int32 / int32 / uint32 numpy VIEWS of pinned
        staging buffers that the next export overwrites (callers map them to real indices at once,
        which copies) -- or ``None`` when the tables are not such a pair or a count does not fit
        24 bits (callers then use :meth:`export_many`)."""
        tables = (atom_table, res_table)
        if any(isinstance(t, HashTable) or not t.is_contiguous() or t.dtype != torch.int32 for t in tables):
            return None
        base = atom_table.untyped_storage().data_ptr()
        n, rc = int(self.n_used), int(self.n_res_c)
        off0, off1 = atom_table.storage_offset(), res_table.storage_offset()
        if (res_table.untyped_storage().data_ptr() != base or atom_table.numel() != n * n or
                res_table.numel() != rc * rc or off0 % 4 or off1 < off0 + n * n or
                off1 + rc * rc - off0 > 2 * (n * n + rc * rc)):
            return None
        span = torch.empty(0, dtype=torch.int32, device=atom_table.device)
        span.set_(atom_table.untyped_storage(), off0, (off1 + rc * rc - off0,))
        hints = self.__dict__.setdefault("_export_hint", {})
        stream = torch.cuda.current_stream(self.device)
        cap = hints.get(span.numel())
        with torch.cuda.device(self.device):
            if cap is None:
                cap = int(torch.count_nonzero(span).item())
            cap = max(16, min(span.numel(), cap + cap // 4 + 1024))
            for attempt in range(2):
                buf = self._export_launch(span, cap, 0, to_host=False)
                if "d_lo" not in buf or buf["d_lo"].numel() < buf["cap"]:
                    for name in ("lo", "hi", "pc"):
                        buf["d_" + name] = torch.empty(buf["cap"], dtype=torch.int32, device=self.device)
                        buf["h_" + name] = torch.empty(buf["cap"], dtype=torch.int32, pin_memory=True)
                    buf["d_split"] = torch.zeros(1, dtype=torch.int64, device=self.device)
                    buf["h_split"] = torch.zeros(1, dtype=torch.int64, pin_memory=True)
                self._check(self._L.cmb_unpack_pairs(self._ctx, _ptr(buf["d_packed"]), _ptr(buf["d_n"]), cap,
                                                     n, off1 - off0, rc, _ptr(buf["d_lo"]), _ptr(buf["d_hi"]),
                                                     _ptr(buf["d_pc"]), _ptr(buf["d_split"]),
                                                     _stream_ptr(self.device)), "cmb_unpack_pairs")
                buf["h_split"].copy_(buf["d_split"], non_blocking=True)
                for name in ("lo", "hi", "pc"):
                    buf["h_" + name][:cap].copy_(buf["d_" + name][:cap], non_blocking=True)
                stream.synchronize()
                found, wide = int(buf["h_n"][0]), int(buf["h_n"][1])
                if wide:
                    return None
                if found <= cap:
                    break
                cap = found                    # capacity guess too small: exact retry
            hints[span.numel()] = found
        k = int(buf["h_split"][0])
        lo, hi = buf["h_lo"][:found].numpy(), buf["h_hi"][:found].numpy()
        cnt = buf["h_pc"][:found].numpy().view(np.uint32)
        return (lo[:k], hi[:k], cnt[:k]), (lo[k:], hi[k:], cnt[k:])

    def export_packed_device(self, tables):
        """Sparse export that STAYS on the device: for every dense table a sorted int64 tensor of
        packed (flat index << 24 | count) words (``cmb_export_sorted``), or ``None`` for a table
        whose counts do not fit 24 bits / a hashed table (callers then use :meth:`export_many`).
        One host synchronisation (the entry counts)."""
        hints = self.__dict__.setdefault("_export_hint", {})
        stream = torch.cuda.current_stream(self.device)
        jobs = []
        with torch.cuda.device(self.device):
            for slot, t in enumerate(tables):
                if isinstance(t, HashTable):
                    jobs.append(None)
                    continue
                cap = hints.get(t.numel())
                if cap is None:
                    cap = int(torch.count_nonzero(t).item())
                cap = max(16, min(t.numel(), cap + cap // 4 + 1024))
                jobs.append(self._export_launch(t, cap, slot, to_host=False))
            stream.synchronize()
            out = []
            for slot, (t, job) in enumerate(zip(tables, jobs)):
                if job is None:
                    out.append(None)
                    continue
                n, wide = int(job["h_n"][0]), int(job["h_n"][1])
                if n > job["use"] and not wide:
                    job = self._export_launch(t, n, slot, to_host=False)
                    stream.synchronize()
                    n, wide = int(job["h_n"][0]), int(job["h_n"][1])
                if wide:
                    out.append(None)
                    continue
                hints[t.numel()] = n
                out.append(job["d_packed"][:n])
        return out

    def _export_span(self, tables):
        """Dense tables that are views of ONE allocation (``new_tables``) are exported with a single
        scan of that allocation and split on the host at the view boundaries."""
        if len(tables) < 2 or any(isinstance(t, HashTable) or not t.is_contiguous() for t in tables):
            return None
        base = tables[0].untyped_storage().data_ptr()
        if any(t.untyped_storage().data_ptr() != base or t.dtype != tables[0].dtype for t in tables):
            return None
        offsets = [t.storage_offset() for t in tables]
        ends = [o + t.numel() for o, t in zip(offsets, tables)]
        if offsets != sorted(offsets) or any(e > o2 for e, o2 in zip(ends[:-1], offsets[1:])):
            return None
        lo, hi = offsets[0], ends[-1]
        if hi - lo > 2 * sum(t.numel() for t in tables) or lo % 4:
            return None
        span = torch.empty(0, dtype=tables[0].dtype, device=tables[0].device)
        span.set_(tables[0].untyped_storage(), lo, (hi - lo,))
        (idx, cnt), = self.export_many([span])
        cuts = np.searchsorted(idx, np.asarray(offsets + [hi], dtype=np.int64) - lo)
        return [(idx[a:b] if o == lo else idx[a:b] - (o - lo), cnt[a:b]) for a, b, o in zip(cuts[:-1], cuts[1:], offsets)]

    def _export_launch(self, table, cap, slot, to_host=True):
        # staging buffers are keyed by (table size, position in the call) and reused across calls
        pool = self.__dict__.setdefault("_export_pool", {})
        buf = pool.get((table.numel(), slot))
        if buf is None or buf["cap"] < cap:
            buf = {
                "cap": cap,
                "d_packed": torch.empty(cap, dtype=torch.int64, device=self.device),
                "d_n": torch.zeros(2, dtype=torch.int64, device=self.device),
                "h_n": torch.zeros(2, dtype=torch.int64, pin_memory=True),
            }
            if len(pool) > 8:
                pool.clear()
            pool[(table.numel(), slot)] = buf
        hashed = isinstance(table, HashTable)
        # the kernels are given the REQUESTED capacity (the pooled buffers may be larger): entries
        # beyond it are counted but not written, and only `use` entries are copied to the host
        buf["use"] = cap
        self._check(self._L.cmb_export_sorted(self._ctx, _table_ptr(table), table.numel(),
                                              table.log2 if hashed else 0, cap, _ptr(buf["d_packed"]),
                                              _ptr(buf["d_n"]), _stream_ptr(self.device)), "cmb_export_sorted")
        buf["h_n"].copy_(buf["d_n"], non_blocking=True)
        if to_host:
            if "d_idx" not in buf:
                buf["d_idx"] = torch.empty(buf["cap"], dtype=torch.int64, device=self.device)
                buf["d_cnt"] = torch.empty(buf["cap"], dtype=torch.int32, device=self.device)
                buf["h_idx"] = torch.empty(buf["cap"], dtype=torch.int64, pin_memory=True)
                buf["h_cnt"] = torch.empty(buf["cap"], dtype=torch.int32, pin_memory=True)
            self._check(self._L.cmb_unpack_sorted(self._ctx, _ptr(buf["d_packed"]), _ptr(buf["d_n"]), cap, 0,
                                                  _ptr(buf["d_idx"]), _ptr(buf["d_cnt"]), _stream_ptr(self.device)),
                        "cmb_unpack_sorted")
            # (the capacity is the previous entry count + 25 %: copying all of it costs less than a
            # second synchronisation to learn the exact count first)
            buf["h_idx"][:cap].copy_(buf["d_idx"][:cap], non_blocking=True)
            buf["h_cnt"][:cap].copy_(buf["d_cnt"][:cap], non_blocking=True)
        return buf

    def _export_unpacked(self, table):
        """Export through ``cmb_export_nonzero`` (counts of 2^24 and more), sorted on the host."""
        n = int(torch.count_nonzero(table).item())
        d_idx = torch.empty(max(n, 1), dtype=torch.int64, device=self.device)
        d_cnt = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
        d_n = torch.zeros(1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._L.cmb_export_nonzero(self._ctx, _ptr(table), table.numel(), max(n, 1), _ptr(d_idx),
                                                   _ptr(d_cnt), _ptr(d_n), _stream_ptr(self.device)),
                        "cmb_export_nonzero")
        idx = d_idx[:n].cpu().numpy()
        cnt = d_cnt[:n].cpu().numpy().view(np.uint32)
        order = np.argsort(idx, kind="stable")
        return idx[order], cnt[order]

    def frame_contacts(self, xyz, box, frame0=0, atom_table=None, res_table=None,
                       want_atoms=True, want_residues=True, cap_hint=None):
        """Per-frame contact keys of a DEVICE trajectory.

        Returns (atom_keys, res_keys): sorted uint64 numpy arrays with
        key = (frame0 + f) << 40 | lo << 20 | hi.
        """
        xyz, box = self._dev_inputs(xyz, box)
        n_frames = xyz.shape[0]
        # capacities: keys per frame seen in the previous call for this system (+25 %), else 64
        hint = self.__dict__.setdefault("_keys_hint", {}).get(self._system_key, (64.0, 64.0))
        cap_a = int(cap_hint or max(1024, n_frames * hint[0] * 1.25 + 1024)) if want_atoms else 0
        cap_r = int(cap_hint or max(1024, n_frames * hint[1] * 1.25 + 1024)) if want_residues else 0
        if isinstance(atom_table, HashTable) or isinstance(res_table, HashTable):
            raise NotImplementedError("frame_contacts accumulates into dense tables only")
        self._set_table_mode(None, None)
        if atom_table is not None or res_table is not None:
            # tables must only be touched once: size the key buffers with a dry run first
            counts = self._frame_contacts_once(xyz, box, frame0, None, 0, None, 0, None, None,
                                               count_only=(want_atoms, want_residues))
            cap_a, cap_r = (int(counts[0]) if want_atoms else 0), (int(counts[1]) if want_residues else 0)
        while True:
            ak = torch.empty(max(cap_a, 1), dtype=torch.int64, device=self.device) if want_atoms else None
            rk = torch.empty(max(cap_r, 1), dtype=torch.int64, device=self.device) if want_residues else None
            counts = self._frame_contacts_once(xyz, box, frame0, ak, cap_a, rk, cap_r,
                                               atom_table, res_table)
            na, nr = int(counts[0]), int(counts[1])
            if (not want_atoms or na <= cap_a) and (not want_residues or nr <= cap_r):
                break
            if atom_table is not None or res_table is not None:
                raise EngineError("key buffers overflowed after sizing pass")
            cap_a, cap_r = (max(cap_a, na) if want_atoms else 0), (max(cap_r, nr) if want_residues else 0)
        repeated = self._system_key is not None and self._system_key in self._keys_hint
        if self._system_key is not None and n_frames > 0:
            self._keys_hint[self._system_key] = (max(1.0, na / n_frames) if want_atoms else hint[0],
                                                 max(1.0, nr / n_frames) if want_residues else hint[1])
        out = []
        for keys, n_keys in ((ak, na), (rk, nr)):
            if keys is None:
                out.append(None)
                continue
            keys = keys[:n_keys]
            if n_keys > 1:
                with torch.cuda.device(self.device):
                    self._check(self._L.cmb_sort_keys(self._ctx, _ptr(keys), n_keys,
                                                      _stream_ptr(self.device)), "cmb_sort_keys")
            # Repeated calls (a trajectory processed chunk by chunk) copy into pinned memory from
            # torch's caching host allocator: the copy then runs at PCIe speed and the block is
            # recycled once the returned array, which keeps it alive, is dropped.  The first call
            # for a system stays pageable: pinning 100 MB costs more than one slow copy.
            host = torch.empty(n_keys, dtype=torch.int64, pin_memory=repeated)
            host.copy_(keys, non_blocking=True)
            out.append(host)
        torch.cuda.current_stream(self.device).synchronize()
        out = [None if h is None else h.numpy().view(np.uint64) for h in out]
        return out[0], out[1]

    def frame_csr(self, xyz, box, atom_map, res_map, used=None, frame0=0):
        """Per-frame contacts of a whole trajectory as CSR arrays, everything but the final copy on
        the device: key emission -> radix sort -> ``cmb_keys_to_csr`` -> pinned host buffers.

        ``xyz``: CUDA tensor ``[F, n_used, 3]``, or a HOST array ``[F, n_total, 3]`` streamed through
        ``cmb_frame_contacts_host`` (``used``: ascending indices of the used atoms when the array
        still holds all atoms).  ``atom_map`` / ``res_map``: compact -> real index.
        Returns ``((ptr, i, j), (ptr, i, j))`` for atoms and residues: ``ptr`` int64 ``[F + 1]``,
        ``i < j`` int32 real indices, rows sorted by (i, j) within a frame.
        """
        on_device = isinstance(xyz, torch.Tensor) and xyz.is_cuda
        if on_device:
            xyz, box = self._dev_inputs(xyz, box)
        else:
            xyz = xyz.numpy() if isinstance(xyz, torch.Tensor) else xyz
            box = box.numpy() if isinstance(box, torch.Tensor) else box
            if xyz.dtype != np.float32 or not xyz.flags["C_CONTIGUOUS"] or xyz.ndim != 3:
                raise TypeError("xyz must be C-contiguous float32 [n_frames, n_atoms, 3]")
            if used is None and xyz.shape[1] != self.n_used:
                raise ValueError("xyz must hold the %d used atoms (or pass `used`)" % self.n_used)
            if used is not None:
                used = np.ascontiguousarray(used, dtype=np.int32)
                if used.shape[0] != self.n_used:
                    raise ValueError("len(used) must equal the number of used atoms of the system")
            if box is not None and (box.dtype != np.float32 or not box.flags["C_CONTIGUOUS"]
                                    or box.size != xyz.shape[0] * 9):
                raise TypeError("box must be C-contiguous float32 [n_frames, 3, 3]")
        n_frames = int(xyz.shape[0])
        self._set_table_mode(None, None)
        dev = self.device
        marks = [("start", time.perf_counter())]
        hints = self.__dict__.setdefault("_keys_hint", {})
        hint = hints.get(self._system_key)
        # Repeated calls for a system copy the result into pinned memory from torch's caching host
        # allocator (PCIe speed; the block is recycled once the returned arrays are dropped).  The
        # first call stays pageable: pinning 80 MB costs several times more than one slower copy.
        repeated = hint is not None
        if hint is None:
            # keys per frame from a counting pass over a short prefix (nothing is written)
            k = min(n_frames, 64)
            if on_device:
                px, pb = xyz[:k], (None if box is None else box[:k])
            else:
                px = np.ascontiguousarray(xyz[:k] if used is None else xyz[:k][:, used])
                px = torch.as_tensor(px, device=dev)
                pb = None if box is None else torch.as_tensor(np.ascontiguousarray(box[:k]), device=dev)
            counts = self._frame_contacts_once(px, pb, 0, None, 0, None, 0, None, None, count_only=(True, True))
            hint = (max(1.0, float(counts[0]) / k), max(1.0, float(counts[1]) / k))
        cap_a = int(n_frames * hint[0] * 1.25 + 4096)
        cap_r = int(n_frames * hint[1] * 1.25 + 4096)
        stream = torch.cuda.current_stream(dev)
        while True:
            ak = torch.empty(cap_a, dtype=torch.int64, device=dev)
            rk = torch.empty(cap_r, dtype=torch.int64, device=dev)
            counts = torch.zeros(4, dtype=torch.int64, device=dev)
            with torch.cuda.device(dev):
                if on_device:
                    self._check(self._L.cmb_frame_contacts(self._ctx, _ptr(xyz), _ptr(box), n_frames, int(frame0),
                                                           _ptr(ak), cap_a, _ptr(rk), cap_r, _ptr(counts), None,
                                                           None, _stream_ptr(dev)), "cmb_frame_contacts")
                else:
                    stream.synchronize()                     # counters zeroed before the engine's streams start
                    self._check(self._L.cmb_frame_contacts_host(
                        self._ctx, _ptr(xyz), int(xyz.shape[1]), _ptr(used), _ptr(box), n_frames, int(frame0), 0, 0,
                        _ptr(ak), cap_a, _ptr(rk), cap_r, _ptr(counts)), "cmb_frame_contacts_host")
            na, nr = (int(v) for v in counts[:2].cpu().numpy())
            marks.append(("emit", time.perf_counter()))
            if na <= cap_a and nr <= cap_r:
                break
            cap_a, cap_r = max(cap_a, na + na // 8), max(cap_r, nr + nr // 8)
        if self._system_key is not None and n_frames > 0:
            hints[self._system_key] = (max(1.0, na / n_frames), max(1.0, nr / n_frames))
        maps = self.__dict__.setdefault("_csr_maps", {})
        key = (self._system_key, id(atom_map), id(res_map))
        if key not in maps:
            maps.clear()
            maps[key] = (torch.as_tensor(np.ascontiguousarray(atom_map, dtype=np.int32), device=dev),
                         torch.as_tensor(np.ascontiguousarray(res_map, dtype=np.int32), device=dev),
                         atom_map, res_map)          # (the arrays are kept alive so that their ids stay theirs)
        d_maps = maps[key]
        out = []
        with torch.cuda.device(dev):
            for keys, n, d_map in ((ak, na, d_maps[0]), (rk, nr, d_maps[1])):
                if n > 1:
                    self._check(self._L.cmb_sort_keys(self._ctx, _ptr(keys), n, _stream_ptr(dev)), "cmb_sort_keys")
                d_ptr = torch.empty(n_frames + 1, dtype=torch.int64, device=dev)
                d_ij = torch.empty((2, max(n, 1)), dtype=torch.int32, device=dev)
                self._check(self._L.cmb_keys_to_csr(self._ctx, _ptr(keys), n, max(n_frames, 1), int(frame0), _ptr(d_map),
                                                    _ptr(d_ptr), _ptr(d_ij[0]), _ptr(d_ij[1]), _stream_ptr(dev)),
                            "cmb_keys_to_csr")
                h_ptr = torch.empty(n_frames + 1, dtype=torch.int64, pin_memory=repeated)
                h_ij = torch.empty((2, max(n, 1)), dtype=torch.int32, pin_memory=repeated)
                h_ptr.copy_(d_ptr, non_blocking=True)
                h_ij.copy_(d_ij, non_blocking=True)
                out.append((h_ptr, h_ij, n))
        marks.append(("launch_sort_csr_copy", time.perf_counter()))
        stream.synchronize()
        marks.append(("sync", time.perf_counter()))
        self.last_timing = {b[0]: b[1] - a[1] for a, b in zip(marks, marks[1:])}
        return tuple((h_ptr.numpy(), h_ij[0, :n].numpy(), h_ij[1, :n].numpy()) for h_ptr, h_ij, n in out)

    def _frame_contacts_once(self, xyz, box, frame0, ak, cap_a, rk, cap_r, atom_table, res_table,
                             count_only=None):
        counts = torch.zeros(4, dtype=torch.int64, device=self.device)
        if count_only is not None:
            # capacity 0 with a non-NULL pointer: keys are counted but never written
            dummy = torch.empty(1, dtype=torch.int64, device=self.device)
            ak = dummy if count_only[0] else None
            rk = dummy if count_only[1] else None
            cap_a = cap_r = 0
        with torch.cuda.device(self.device):
            self._check(self._L.cmb_frame_contacts(self._ctx, _ptr(xyz), _ptr(box), xyz.shape[0],
                                                   int(frame0), _ptr(ak), int(cap_a), _ptr(rk),
                                                   int(cap_r), _ptr(counts), _ptr(atom_table),
                                                   _ptr(res_table), _stream_ptr(self.device)),
                        "cmb_frame_contacts")
        return counts.cpu().numpy()

    def boundary_pairs(self, xyz, box, ulps=1, frame0=0, cap=1 << 16):
        """Eligible pairs whose d2 is within `ulps` float steps of cutoff^2.

        Returns a structured numpy array with fields frame, a, b (int32) and d2 (float32).
        """
        xyz, box = self._dev_inputs(xyz, box)
        self._set_table_mode(None, None)
        dt = np.dtype([("frame", np.int32), ("a", np.int32), ("b", np.int32), ("d2", np.float32)])
        while True:
            rec = torch.zeros(cap * 4, dtype=torch.int32, device=self.device)
            n_out = torch.zeros(1, dtype=torch.int64, device=self.device)
            with torch.cuda.device(self.device):
                self._check(self._L.cmb_boundary_pairs(self._ctx, _ptr(xyz), _ptr(box), xyz.shape[0],
                                                       int(frame0), int(ulps), _ptr(rec), cap,
                                                       _ptr(n_out), _stream_ptr(self.device)),
                            "cmb_boundary_pairs")
            n = int(n_out.item())
            if n <= cap:
                break
            cap = n
        out = rec.cpu().numpy()[:n * 4].view(dt)
        return np.sort(out, order=["frame", "a", "b"])

    def screen_audit(self, xyz, box):
        """(pairs audited, disagreements) of the cell kernels' screening shortcut on a DEVICE
        trajectory: every stencil candidate that is not residue-excluded is decided by the reference
        arithmetic and by the production thresholds; the second number must be 0."""
        xyz, box = self._dev_inputs(xyz, box)
        self._set_table_mode(None, None)
        out = torch.zeros(2, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._L.cmb_screen_audit(self._ctx, _ptr(xyz), _ptr(box), xyz.shape[0], _ptr(out),
                                                 _stream_ptr(self.device)), "cmb_screen_audit")
        audited, bad = out.cpu().numpy().tolist()
        return int(audited), int(bad)

    def gather_atoms(self, xyz_full, used):
        """Device ``_atom_slice``: [F, n_total, 3] -> [F, len(used), 3]."""
        if xyz_full.dtype != torch.float32 or not xyz_full.is_contiguous() or not xyz_full.is_cuda:
            raise TypeError("xyz_full must be a contiguous float32 CUDA tensor")
        used_t = torch.as_tensor(np.ascontiguousarray(used, dtype=np.int32), device=self.device)
        out = torch.empty((xyz_full.shape[0], used_t.numel(), 3), dtype=torch.float32,
                          device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._L.cmb_gather_atoms(self._ctx, _ptr(xyz_full), xyz_full.shape[0],
                                                 xyz_full.shape[1], _ptr(used_t), used_t.numel(),
                                                 _ptr(out), _stream_ptr(self.device)),
                        "cmb_gather_atoms")
        return out

    def min_dist(self, xyz, box, query, haystack, orthogonal):
        """Per-frame minimum distance and first arg-min pair index (query-major product)."""
        if xyz.dtype != torch.float32 or not xyz.is_contiguous() or not xyz.is_cuda:
            raise TypeError("xyz must be a contiguous float32 CUDA tensor")
        q = torch.as_tensor(np.ascontiguousarray(query, dtype=np.int32), device=self.device)
        h = torch.as_tensor(np.ascontiguousarray(haystack, dtype=np.int32), device=self.device)
        n_frames, n_atoms = xyz.shape[0], xyz.shape[1]
        dmin = torch.empty(n_frames, dtype=torch.float32, device=self.device)
        darg = torch.empty(n_frames, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._L.cmb_min_dist(self._ctx, _ptr(xyz), _ptr(box), n_frames, n_atoms,
                                             _ptr(q), q.numel(), _ptr(h), h.numel(),
                                             int(bool(orthogonal)), _ptr(dmin), _ptr(darg),
                                             _stream_ptr(self.device)), "cmb_min_dist")
        return dmin, darg

    def min_dist_pairs(self, xyz, box, atom_pairs, orthogonal):
        """Per-frame minimum distance and first arg-min position over an arbitrary pair list."""
        if xyz.dtype != torch.float32 or not xyz.is_contiguous() or not xyz.is_cuda:
            raise TypeError("xyz must be a contiguous float32 CUDA tensor")
        pairs = torch.as_tensor(np.ascontiguousarray(atom_pairs, dtype=np.int32).reshape(-1, 2), device=self.device)
        n_frames, n_atoms = xyz.shape[0], xyz.shape[1]
        dmin = torch.empty(n_frames, dtype=torch.float32, device=self.device)
        darg = torch.empty(n_frames, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._L.cmb_min_dist_pairs(self._ctx, _ptr(xyz), _ptr(box), n_frames, n_atoms,
                                                   _ptr(pairs), pairs.shape[0], int(bool(orthogonal)),
                                                   _ptr(dmin), _ptr(darg), _stream_ptr(self.device)),
                        "cmb_min_dist_pairs")
        return dmin, darg

    def group_contacts(self, xyz, box, atom_pairs, group_ptr, cutoff, orthogonal):
        """``[n_groups, n_frames]`` bool array: is the minimum ``compute_distances`` distance over
        the atom pairs of each group below ``cutoff`` (float32, strict) in each frame?"""
        if xyz.dtype != torch.float32 or not xyz.is_contiguous() or not xyz.is_cuda:
            raise TypeError("xyz must be a contiguous float32 CUDA tensor")
        pairs = np.ascontiguousarray(atom_pairs, dtype=np.int64).reshape(-1, 2)
        group_ptr = np.ascontiguousarray(group_ptr, dtype=np.int64)
        n_groups, n_frames, n_atoms = len(group_ptr) - 1, xyz.shape[0], xyz.shape[1]
        if n_groups == 0 or n_frames == 0:
            return np.zeros((n_groups, n_frames), dtype=bool)
        if pairs.size and (pairs.min() < 0 or pairs.max() >= n_atoms):
            raise ValueError("atom index out of range")
        atoms, compact = np.unique(pairs, return_inverse=True)
        if atoms.size == 0:
            atoms = np.zeros(1, dtype=np.int64)
        d_atoms = torch.as_tensor(atoms.astype(np.int32), device=self.device)
        d_pairs = torch.as_tensor(compact.reshape(-1, 2).astype(np.int32), device=self.device)
        d_ptr = torch.as_tensor(group_ptr, device=self.device)
        out = torch.empty((n_frames, n_groups), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._L.cmb_group_contacts(self._ctx, _ptr(xyz), _ptr(box), n_frames, n_atoms,
                                                   _ptr(d_atoms), int(d_atoms.numel()), _ptr(d_pairs), _ptr(d_ptr),
                                                   n_groups, int(bool(orthogonal)),
                                                   ctypes.c_float(float(np.float32(cutoff))), _ptr(out),
                                                   _stream_ptr(self.device)), "cmb_group_contacts")
        return out.t().contiguous().cpu().numpy().astype(bool)

    def nearest(self, xyz_frame, box9, cutoff, orthogonal, excl_mode, atom_res=None,
                excl_ptr=None, excl_idx=None):
        n_atoms = xyz_frame.shape[0]
        nearest = torch.empty(n_atoms, dtype=torch.int32, device=self.device)
        dist = torch.empty(n_atoms, dtype=torch.float32, device=self.device)
        keep = []
        def dev(a, dtype):
            if a is None:
                return None
            t = torch.as_tensor(np.ascontiguousarray(a, dtype=dtype), device=self.device)
            keep.append(t)
            return t
        res_t, ptr_t, idx_t = dev(atom_res, np.int32), dev(excl_ptr, np.int64), dev(excl_idx, np.int32)
        with torch.cuda.device(self.device):
            self._check(self._L.cmb_nearest(self._ctx, _ptr(xyz_frame), _ptr(box9), n_atoms,
                                            float(cutoff), int(bool(orthogonal)), int(excl_mode),
                                            _ptr(res_t), _ptr(ptr_t), _ptr(idx_t), _ptr(nearest),
                                            _ptr(dist), _stream_ptr(self.device)), "cmb_nearest")
        torch.cuda.current_stream(self.device).synchronize()
        return nearest, dist

    def synth_frames(self, base, groups, box9, seed, amplitude, image_shifts, frame0, n_frames):
        """Device synthetic generator; bit-identical to ``synth.make_frames``."""
        n_atoms = base.shape[0]
        out = torch.empty((n_frames, n_atoms, 3), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self._L.cmb_synth_frames(self._ctx, _ptr(base), _ptr(groups), n_atoms,
                                                 _ptr(box9), int(seed), float(amplitude),
                                                 int(bool(image_shifts)), int(frame0), int(n_frames),
                                                 _ptr(out), _stream_ptr(self.device)),
                        "cmb_synth_frames")
        return out


_ENGINES = {}


def get_engine(device=None):
    """Process-wide engine cache (one context per device)."""
    if not torch.cuda.is_available():
        raise EngineError("contact_map_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    key = int(device.index if isinstance(device, torch.device) else device)
    if key not in _ENGINES:
        _ENGINES[key] = ContactEngine(key)
    return _ENGINES[key]
