"""CPU, world_size 2 over gloo: the multi-rank frame sharding + single all-reduce of the tables.

The per-rank partial tables are produced by the ORACLE here (no GPU in this container); what is
under test is the product's partition (`parallel.rank_slices`) and reduce (`parallel.reduce_tables`,
`reduce_frame_count`, `gather_keys`).
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from contact_map_b200 import parallel, synth
        from oracle import clib
        from tests.helpers import system_arrays
        rng = np.random.default_rng(11)
        system = synth.random_system(rng, 80, n_chains=2, box_kind="ortho", extent=1.2)
        n_frames = 7
        xyz = system.frames(0, n_frames)
        box = np.broadcast_to(system.box, (n_frames, 3, 3)).copy()
        res, chain, role = system_arrays(system)
        n_res = system.topology.n_residues
        atom = np.zeros((80, 80), dtype=np.uint32)
        resc = np.zeros((n_res, n_res), dtype=np.uint32)
        mine = 0
        keys = []
        for s in parallel.rank_slices(n_frames, n_blocks=3):
            a, r = clib.accumulate_dense(xyz[s], box[s], 0.45, res, chain, role, 2, n_res)
            atom += a
            resc += r
            mine += s.stop - s.start
            for f in range(s.start, s.stop):
                pairs = clib.contacts_frame(xyz[f], box[f], 0.45, res, chain, role, 2)
                keys.append((np.uint64(f) << np.uint64(40)) | (pairs[:, 0].astype(np.uint64) << np.uint64(20))
                            | pairs[:, 1].astype(np.uint64))
        t_atom = torch.from_numpy(atom.view(np.int32).reshape(-1).copy())
        t_res = torch.from_numpy(resc.view(np.int32).reshape(-1).copy())
        # reduce to rank 0 only (the reference's client-side reduce), on two views of one allocation
        both = torch.cat([t_atom, t_res])
        r_atom, r_res = both[:t_atom.numel()], both[t_atom.numel():]
        parallel.reduce_tables([r_atom, r_res], dst=0)
        parallel.reduce_tables([t_atom, t_res])
        total_frames = parallel.reduce_frame_count(mine, torch.device("cpu"))
        all_keys = parallel.gather_keys(np.sort(np.concatenate(keys)) if keys else np.empty(0, np.uint64),
                                        torch.device("cpu"))
        # hashed tables are reduced through their sparse exports (parallel.merge_sparse)
        flat = np.flatnonzero(atom.reshape(-1))
        m_idx, m_cnt = parallel.merge_sparse(flat, atom.reshape(-1)[flat], torch.device("cpu"))
        # ... and, for large dense tables, through packed device-side lists (parallel.merge_packed)
        packed = torch.from_numpy(((flat.astype(np.int64) << 24) | atom.reshape(-1)[flat].astype(np.int64)))
        p_idx, p_cnt = parallel.merge_packed(packed)
        assert parallel.merge_packed(packed, receives=False) is None
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), atom=t_atom.numpy(), res=t_res.numpy(),
                 frames=total_frames, keys=all_keys.view(np.int64), m_idx=m_idx, m_cnt=m_cnt, p_idx=p_idx, p_cnt=p_cnt,
                 root_atom=r_atom.numpy(), root_res=r_res.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_reduce_equals_single(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from contact_map_b200 import synth
    from oracle import clib
    from tests.helpers import system_arrays
    rng = np.random.default_rng(11)
    system = synth.random_system(rng, 80, n_chains=2, box_kind="ortho", extent=1.2)
    xyz = system.frames(0, 7)
    box = np.broadcast_to(system.box, (7, 3, 3)).copy()
    res, chain, role = system_arrays(system)
    atom, resc = clib.accumulate_dense(xyz, box, 0.45, res, chain, role, 2, system.topology.n_residues)
    want_keys = []
    for f in range(7):
        pairs = clib.contacts_frame(xyz[f], box[f], 0.45, res, chain, role, 2)
        want_keys.append((np.uint64(f) << np.uint64(40)) | (pairs[:, 0].astype(np.uint64) << np.uint64(20))
                         | pairs[:, 1].astype(np.uint64))
    want_keys = np.concatenate(want_keys)
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        assert np.array_equal(got["atom"].view(np.uint32).reshape(80, 80), atom)
        assert np.array_equal(got["res"].view(np.uint32).reshape(resc.shape), resc)
        assert int(got["frames"]) == 7
        if rank == 0:
            assert np.array_equal(got["root_atom"].view(np.uint32).reshape(80, 80), atom)
            assert np.array_equal(got["root_res"].view(np.uint32).reshape(resc.shape), resc)
        assert np.array_equal(got["keys"].view(np.uint64), want_keys)
        want_flat = np.flatnonzero(atom.reshape(-1))
        assert np.array_equal(got["m_idx"], want_flat)
        assert np.array_equal(got["m_cnt"], atom.reshape(-1)[want_flat])
        assert np.array_equal(got["p_idx"], want_flat)
        assert np.array_equal(got["p_cnt"], atom.reshape(-1)[want_flat])
    assert atom.sum() > 0


def _file_worker(rank, world, port, out_dir, file_name):
    """Each rank maps the trajectory FILE and touches only the frames of its own blocks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from contact_map_b200 import frequency_task, parallel
        from oracle import clib
        from tests.helpers import system_arrays_of_topology
        probe = frequency_task.load_trajectory_task(slice(0, 0), file_name)      # topology only, no frames read
        res, chain, role = system_arrays_of_topology(probe.topology)
        n, n_res = probe.topology.n_atoms, probe.topology.n_residues
        n_frames = len(frequency_task.load_trajectory_task(slice(None), file_name))
        atom = np.zeros((n, n), dtype=np.uint32)
        resc = np.zeros((n_res, n_res), dtype=np.uint32)
        touched = []
        for s in parallel.rank_slices(n_frames, n_blocks=5):
            part = frequency_task.load_trajectory_task(s, file_name)
            assert not part.xyz.flags["OWNDATA"]                      # still a view of the mapped file
            a, r = clib.accumulate_dense(np.ascontiguousarray(part.xyz), np.ascontiguousarray(part.unitcell_vectors),
                                         0.45, res, chain, role, 2, n_res)
            atom += a
            resc += r
            touched.extend(range(s.start, s.stop))
        t_atom = torch.from_numpy(atom.view(np.int32).reshape(-1).copy())
        t_res = torch.from_numpy(resc.view(np.int32).reshape(-1).copy())
        parallel.reduce_tables([t_atom, t_res])
        total = parallel.reduce_frame_count(len(touched), torch.device("cpu"))
        np.savez(os.path.join(out_dir, "file_rank%d.npz" % rank), atom=t_atom.numpy(), res=t_res.numpy(),
                 frames=total, touched=np.asarray(touched, dtype=np.int64))
    finally:
        dist.destroy_process_group()


def test_two_ranks_read_disjoint_slices_of_a_mapped_file(tmp_path):
    """f2: the chunked reader sharded over two ranks -- every frame of the file is read by exactly
    one rank and the reduced tables equal the in-memory single-process run."""
    from contact_map_b200 import synth, trajectory
    from oracle import clib
    from tests.helpers import system_arrays
    rng = np.random.default_rng(5)
    system = synth.random_system(rng, 90, n_chains=2, box_kind="triclinic", extent=1.6)
    n_frames = 23
    xyz = system.frames(0, n_frames)
    box = np.broadcast_to(system.box, (n_frames, 3, 3)).copy()
    file_name = str(tmp_path / "sharded.npy")
    trajectory.save_npy(file_name, trajectory.Trajectory(xyz, system.topology, box))
    port = _free_port()
    mp.spawn(_file_worker, args=(2, port, str(tmp_path), file_name), nprocs=2, join=True)
    res, chain, role = system_arrays(system)
    atom, resc = clib.accumulate_dense(xyz, box, 0.45, res, chain, role, 2, system.topology.n_residues)
    seen = []
    for rank in range(2):
        got = np.load(os.path.join(str(tmp_path), "file_rank%d.npz" % rank))
        assert np.array_equal(got["atom"].view(np.uint32).reshape(atom.shape), atom)
        assert np.array_equal(got["res"].view(np.uint32).reshape(resc.shape), resc)
        assert int(got["frames"]) == n_frames
        seen.extend(got["touched"].tolist())
    assert sorted(seen) == list(range(n_frames))
    assert atom.sum() > 0
