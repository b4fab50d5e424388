"""Frame-chunked runners: the GPU replacement of the reference's Dask runner.

Mirror of /root/reference/contact_map/dask_runner.py (``dask_run`` :11-37,
``DaskContactFrequency`` :66-132, ``DaskContactTrajectory`` :135-215).  The reference maps
frame blocks to Dask worker processes and reduces pickled ``ContactFrequency`` objects on one
worker.  Here the blocks of the SAME partition (``frequency_task.default_slices``) are dealt to
the ranks of a ``torch.distributed`` job (one process per GPU); each rank streams its blocks
through the engine into one pair of device tables and a single all-reduce sums them.
With one process it is simply a chunked single-GPU run.
"""
import numpy as np

from . import parallel
from . import trajectory as _trajectory
from .contact_map import (ContactFrequency, ContactObject, _engine_system, _split_trajectory,
                          _exports_to_pairs, _table_kind, _tables_to_pairs, _used_coordinates)
from .contact_count import PairTable
from .contact_trajectory import KEY_MAX_FRAMES, KEY_MAX_INDEX, ContactTrajectory, FrameCSR, _build_contacts


def _n_workers(client):
    if client is None:
        return None
    ncores = getattr(client, "ncores", None)
    if callable(ncores):
        return len(ncores())
    return int(client)


class ChunkedContactFrequency(ContactFrequency):
    """ContactFrequency computed block by block, blocks sharded over the ranks.

    ``n_blocks`` plays the role of the reference's worker count in
    ``default_slices(len(trajectory), n_workers)`` (dask_runner.py:58-59).
    """

    def __init__(self, trajectory, query=None, haystack=None, cutoff=0.45, n_neighbors_ignored=2,
                 n_blocks=None, presharded=False, result_rank=None):
        """``presharded=True``: ``trajectory`` already holds only THIS rank's frames (each rank
        loaded / generated its own shard); every local frame is processed and ``n_frames``
        becomes the sum over the ranks.  ``result_rank=r``: the partial tables are reduced to
        rank ``r`` only, as the reference's client-side ``reduce_all_results`` does (half the
        traffic of the default all-reduce); the other ranks end up with empty contact maps."""
        self._n_blocks = n_blocks
        self._presharded = presharded
        self._result_rank = result_rank
        super(ChunkedContactFrequency, self).__init__(trajectory, query, haystack, cutoff,
                                                      n_neighbors_ignored)

    def _build_contact_map(self, trajectory):
        from .engine import get_engine
        engine = get_engine()
        used = _engine_system(self, engine)
        xyz, box = _split_trajectory(trajectory)
        # (host arrays keep all their atoms: the used ones are gathered block by block into pinned
        # staging, so a memory-mapped trajectory is only ever read where this rank's blocks lie)
        xyz, box, on_device = _used_coordinates(xyz, box, used, self.topology.n_atoms, engine,
                                                keep_full_host=True)
        atom_table, res_table = engine.new_tables()
        if self._presharded:
            self.block_slices = [slice(0, xyz.shape[0])]
            self._n_frames = parallel.reduce_frame_count(xyz.shape[0], engine.device)
        else:
            self.block_slices = parallel.rank_slices(xyz.shape[0], self._n_blocks)
        for block in self.block_slices:
            bx = xyz[block]
            bb = None if box is None else box[block]
            if bx.shape[0] == 0:
                continue
            if on_device:
                engine.accumulate(bx, bb, atom_table, res_table)
            elif bx.shape[1] != used.shape[0]:
                engine.accumulate_host_gather(bx, used, bb, atom_table, res_table)
            else:
                engine.accumulate_host(bx, bb, atom_table, res_table)
        self._table_kind = _table_kind(atom_table)
        if parallel.world()[1] == 1:
            return _tables_to_pairs(engine, atom_table, res_table, used)
        exports = parallel.reduce_export(engine, [atom_table, res_table], dst=self._result_rank)
        if exports is None:
            return PairTable.empty(), PairTable.empty()
        return _exports_to_pairs(engine, exports, used)


class DaskContactFrequency(ChunkedContactFrequency):
    """Drop-in for the reference's ``DaskContactFrequency(client, filename, ...)``.

    ``client`` is only asked for its worker count (``len(client.ncores())``) to reproduce the
    reference partition; it may be ``None`` or an int.  ``filename`` is loaded with
    ``trajectory.load`` (``.pdb`` / ``.npz``, or a memory-mapped ``.npy`` of which every rank then
    reads only its own frame blocks); a trajectory object is accepted as well.
    """

    def __init__(self, client, filename, query=None, haystack=None, cutoff=0.45,
                 n_neighbors_ignored=2, **kwargs):
        self.client = client
        self.filename = filename
        self.kwargs = kwargs
        trajectory = _trajectory.load(filename, **kwargs) if isinstance(filename, str) else filename
        super(DaskContactFrequency, self).__init__(trajectory, query, haystack, cutoff,
                                                   n_neighbors_ignored, n_blocks=_n_workers(client))

    @property
    def parameters(self):
        return {"query": self.query, "haystack": self.haystack, "cutoff": self.cutoff,
                "n_neighbors_ignored": self.n_neighbors_ignored}

    @property
    def run_info(self):
        return {"parameters": self.parameters, "trajectory_file": self.filename,
                "load_kwargs": self.kwargs}


class DaskContactTrajectory(ContactTrajectory):
    """Drop-in for the reference's ``DaskContactTrajectory(client, filename, ...)``: blocks are
    sharded over the ranks and the per-rank key streams concatenated in frame order."""

    def __init__(self, client, filename, query=None, haystack=None, cutoff=0.45,
                 n_neighbors_ignored=2, **kwargs):
        self.client = client
        self.filename = filename
        self.kwargs = kwargs
        self.trajectory = _trajectory.load(filename, **kwargs) if isinstance(filename, str) else filename
        self._n_blocks = _n_workers(client)
        super(DaskContactTrajectory, self).__init__(self.trajectory, query, haystack, cutoff,
                                                    n_neighbors_ignored)

    def _build_contacts(self, trajectory):
        from .engine import get_engine
        rank, world_size = parallel.world()
        n_frames = len(trajectory)
        if world_size == 1 and self._n_blocks in (None, 1):
            return _build_contacts(self, trajectory)
        engine = get_engine()
        if n_frames > KEY_MAX_FRAMES:
            raise ValueError("DaskContactTrajectory handles at most %d frames (24-bit frame ids in the "
                             "per-frame keys)" % KEY_MAX_FRAMES)
        # keys travel between the ranks in COMPACT numbering (rank among the used atoms / used
        # residues, < 2^20 by the C limit) and are mapped to real indices after the gather, so a
        # topology with more than 2^20 atoms and a small query still packs correctly
        used = np.asarray(self._all_arr, dtype=np.int64)
        atom_res, _ = _trajectory.topology_arrays(self.topology)
        used_res = np.unique(atom_res[used]).astype(np.int64)
        if used.shape[0] > KEY_MAX_INDEX or used_res.shape[0] > KEY_MAX_INDEX:
            raise ValueError("per-frame keys hold 20-bit indices: at most %d used atoms" % KEY_MAX_INDEX)
        atom_keys, res_keys = [], []
        for block in parallel.rank_slices(n_frames, self._n_blocks):
            atoms, residues = _build_contacts(self, trajectory[block])
            shift = np.uint64(block.start) << np.uint64(40)
            for csr, sink, compact in ((atoms, atom_keys, used), (residues, res_keys, used_res)):
                frame = np.repeat(np.arange(csr.n_frames, dtype=np.uint64), np.diff(csr.ptr))
                lo = np.searchsorted(compact, csr.i).astype(np.uint64)
                hi = np.searchsorted(compact, csr.j).astype(np.uint64)
                sink.append(((frame << np.uint64(40)) + shift) | (lo << np.uint64(20)) | hi)
        cat = lambda parts: np.sort(np.concatenate(parts)) if parts else np.empty(0, np.uint64)  # noqa: E731
        atom_keys = parallel.gather_keys(cat(atom_keys), engine.device)
        res_keys = parallel.gather_keys(cat(res_keys), engine.device)
        return (FrameCSR.from_keys(n_frames, atom_keys, used),
                FrameCSR.from_keys(n_frames, res_keys, used_res))
