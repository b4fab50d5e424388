"""GPU x2: the multi-GPU reduce of the C ABI (`cmb_reduce`) driven WITHOUT torch --
two processes, one per GPU, each a plain ctypes host: context, tables from `cmb_device_alloc`, its
shard of the frames through `cmb_accumulate_host_gather`, a NCCL communicator from
`cmb_nccl_unique_id` / `cmb_nccl_comm_create`, ONE in-place collective over the table allocation
(all-reduce, then a rooted reduce), tables copied back and compared with the single-process oracle.
Replaces reduce_all_results (/root/reference/contact_map/frequency_task.py:125-141).
Skipped on boxes with fewer than two GPUs (run with `gpurun --gpus 2`)."""
import ctypes
import multiprocessing
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, tmp, n_frames):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from contact_map_b200 import _lib, frequency_task, synth
    L = _lib.lib()
    ctx = ctypes.c_void_p()
    assert L.cmb_ctx_create(rank, ctypes.byref(ctx)) == 0

    def ck(rc, what):
        assert rc == 0, "%s: %s" % (what, L.cmb_last_error(ctx))
    rng = np.random.default_rng(3)
    system = synth.random_system(rng, 400, n_chains=2, box_kind="triclinic", extent=2.2)
    xyz = system.frames(0, n_frames)
    box = np.broadcast_to(system.box, (n_frames, 3, 3)).copy()
    res, chain = system.topology.atom_residue_chain()
    n = system.n_atoms
    role = np.full(n, 3, dtype=np.uint8)
    res, chain = res.astype(np.int32), chain.astype(np.int32)
    ck(L.cmb_set_system(ctx, n, res.ctypes.data_as(ctypes.c_void_p), chain.ctypes.data_as(ctypes.c_void_p),
                        role.ctypes.data_as(ctypes.c_void_p), 0.45, 2), "set_system")
    a_el, r_el = ctypes.c_int64(0), ctypes.c_int64(0)
    ck(L.cmb_table_sizes(ctx, ctypes.byref(a_el), ctypes.byref(r_el)), "table_sizes")
    a_pad = (a_el.value + 3) // 4 * 4
    total = a_pad + r_el.value
    tables = ctypes.c_void_p()
    ck(L.cmb_device_alloc(ctx, 4 * total, ctypes.byref(tables)), "device_alloc")
    res_tab = ctypes.c_void_p(tables.value + 4 * a_pad)
    used = np.arange(n, dtype=np.int32)
    # communicator: rank 0 creates the id, the others read its 128 bytes from a file
    id_file = os.path.join(tmp, "nccl_id.bin")
    nccl_id = (ctypes.c_char * 128)()
    if rank == 0:
        ck(L.cmb_nccl_unique_id(ctx, nccl_id), "nccl_unique_id")
        with open(id_file + ".tmp", "wb") as fh:
            fh.write(bytes(nccl_id))
        os.replace(id_file + ".tmp", id_file)
    else:
        for _ in range(600):
            if os.path.exists(id_file):
                break
            time.sleep(0.05)
        ctypes.memmove(nccl_id, open(id_file, "rb").read(), 128)
    comm = ctypes.c_void_p()
    ck(L.cmb_nccl_comm_create(ctx, world, nccl_id, rank, ctypes.byref(comm)), "nccl_comm_create")
    host = np.empty(total, dtype=np.uint32)
    for root in (-1, 0):
        ck(L.cmb_device_zero(ctx, tables, 4 * total, None), "device_zero")
        for k, sl in enumerate(frequency_task.default_slices(n_frames, 5)):
            if k % world != rank:
                continue
            part, pbox = np.ascontiguousarray(xyz[sl]), np.ascontiguousarray(box[sl])
            ck(L.cmb_accumulate_host_gather(ctx, part.ctypes.data_as(ctypes.c_void_p), n,
                                            used.ctypes.data_as(ctypes.c_void_p), pbox.ctypes.data_as(ctypes.c_void_p),
                                            len(part), 0, 0, tables, res_tab), "accumulate")
        ck(L.cmb_reduce(ctx, comm, tables, total, root, None), "cmb_reduce")
        ck(L.cmb_copy_to_host(ctx, host.ctypes.data_as(ctypes.c_void_p), tables, 4 * total, None), "copy_to_host")
        np.save(os.path.join(tmp, "rank%d_root%d.npy" % (rank, root)), host)
    ck(L.cmb_nccl_comm_destroy(ctx, comm), "comm_destroy")
    ck(L.cmb_device_free(ctx, tables), "device_free")
    L.cmb_ctx_destroy(ctx)


def test_cmb_reduce_two_ranks_equal_the_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from contact_map_b200 import synth
    from oracle import clib
    n_frames = 17
    mp = multiprocessing.get_context("spawn")
    procs = [mp.Process(target=_worker, args=(r, 2, str(tmp_path), n_frames)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    rng = np.random.default_rng(3)
    system = synth.random_system(rng, 400, n_chains=2, box_kind="triclinic", extent=2.2)
    xyz = system.frames(0, n_frames)
    box = np.broadcast_to(system.box, (n_frames, 3, 3)).copy()
    res, chain = system.topology.atom_residue_chain()
    n = system.n_atoms
    n_res = system.topology.n_residues
    want_a, want_r = clib.accumulate_dense(xyz, box, 0.45, res.astype(np.int32), chain.astype(np.int32),
                                           np.full(n, 3, np.uint8), 2, n_res)
    a_pad = (n * n + 3) // 4 * 4
    for root, ranks in ((-1, (0, 1)), (0, (0,))):          # all-reduce: both ranks; rooted: rank 0 only
        for rank in ranks:
            got = np.load(os.path.join(str(tmp_path), "rank%d_root%d.npy" % (rank, root)))
            assert np.array_equal(got[:n * n].reshape(n, n), want_a)
            assert np.array_equal(got[a_pad:a_pad + n_res * n_res].reshape(n_res, n_res), want_r)
    assert want_a.sum() > 0
