/*
 * cmb200.h -- C ABI of libcmb200.so, the B200 (sm_100a) contact-map engine.
 *
 * Plain pointers and sizes only; no torch / C++ types.  Device pointers are raw
 * CUDA device addresses (e.g. torch.Tensor.data_ptr()); `stream` is a
 * cudaStream_t passed as void* (0 = default stream).  All functions return 0 on
 * success or a negative error code; cmb_last_error() gives the message.  The
 * library never exits/aborts and has no CPU fallback: without a CUDA device every
 * compute entry point fails with CMB_ERR_CUDA.
 *
 * The reference (dwhswenson/contact_map) has no FFI of its own: its only native
 * boundary is MDTraj's Cython API, and its extension hooks are the two private
 * methods its Dask runner overrides.  Each entry point below names the reference
 * interface (file:line under /root/reference/contact_map/) it replaces.
 */
#ifndef CMB200_H
#define CMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMB_OK 0
#define CMB_ERR_ARG (-1)      /* bad argument */
#define CMB_ERR_CUDA (-2)     /* CUDA runtime error (message has the detail) */
#define CMB_ERR_STATE (-3)    /* call order violated (e.g. no system set) */
#define CMB_ERR_LIMIT (-4)    /* system exceeds what this build supports */

typedef struct cmb_ctx cmb_ctx;

int cmb_version(void);
const char* cmb_last_error(const cmb_ctx* ctx);   /* ctx may be NULL: last create error */

/* One context per (process, device); not thread safe. */
int cmb_ctx_create(int device, cmb_ctx** out);
void cmb_ctx_destroy(cmb_ctx* ctx);

/*
 * Describe the system (host pointers, copied).  Replaces the per-object set-up of
 * ContactObject.__init__ (contact_map.py:120-150), residue_neighborhood /
 * _residue_ignore_atom_idxs (contact_map.py:41-60,442-461) and the indexer tables
 * (atom_indexer.py:32-106) -- all collapsed into per-used-atom arrays:
 *   atom_res[i]    global residue index of used atom i
 *   atom_chain[i]  chain index of that residue
 *   role[i]        bit0 = in query, bit1 = in haystack
 * A pair {a,b} is eligible iff (a in Q and b in H) or (b in Q and a in H), and not
 * (chain equal and |res_a - res_b| <= n_ignored).  cutoff is narrowed to float
 * exactly as MDTraj does.  n_ignored must be >= 0.
 */
int cmb_set_system(cmb_ctx* ctx, int n_used, const int32_t* atom_res, const int32_t* atom_chain,
                   const uint8_t* role, double cutoff, int n_ignored);

/* Number of distinct residues among the used atoms (R_c) and their global indices. */
int cmb_n_residues(const cmb_ctx* ctx);
int cmb_residue_map(const cmb_ctx* ctx, int32_t* out_global_res /* [R_c] */);
/* Table sizes in uint32 elements: atoms n_used*n_used (entry [lo*n+hi], lo<hi used
 * indices), residues R_c*R_c (entry [rlo*R_c+rhi], compact residue ids). */
int cmb_table_sizes(const cmb_ctx* ctx, int64_t* atom_elems, int64_t* res_elems);
/* 0 = fused shared-memory path (n_used <= 6144), 1 = global-memory cell path. */
int cmb_path_kind(const cmb_ctx* ctx);

/*
 * ContactFrequency._build_contact_map (contact_map.py:717-744): add n_frames frames
 * into caller-owned DEVICE tables (uint32, zero-initialised by the caller).
 *   d_xyz  [n_frames][n_used][3] fp32 (already sliced to the used atoms, see
 *          cmb_gather_atoms);  d_box [n_frames][9] fp32 row vectors, or NULL for a
 *          trajectory without unit cell.
 * Asynchronous on `stream`.
 */
int cmb_accumulate(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames,
                   uint32_t* d_atom_table, uint32_t* d_res_table, void* stream);

/*
 * Same, with HOST coordinates: frames are staged through two device buffers, host->device
 * copies overlapping the kernels.  Pinned memory is copied directly; pageable memory (an
 * ordinary array, a memory-mapped file) is first packed by host threads into pinned staging
 * buffers, chunk k + 1 while chunk k is transferred and processed (2.3x faster than letting the
 * driver stage it; cmb_set_option "stage_pageable" 0 switches that off).
 * chunk_frames <= 0 lets the library choose the schedule: chunks of about 80 MiB (at least four per
 * call), a short first chunk (the pair list is built from it while the second chunk is copied) and a
 * quarter-size last chunk (the call ends one chunk's worth of kernels after the last copy).
 * Synchronous.
 */
int cmb_accumulate_host(cmb_ctx* ctx, const float* h_xyz, const float* h_box, int64_t n_frames,
                        int64_t chunk_frames, uint32_t* d_atom_table, uint32_t* d_res_table);

/*
 * The host entry points run on the context's own streams.  A caller that has queued work on ITS
 * stream which the next host call depends on (typically: zero-filling the count tables) either
 * synchronises that stream first or calls this: the kernels of the next cmb_accumulate_host /
 * cmb_accumulate_host_gather call wait for everything queued on `stream` so far, its first
 * host-to-device copies do not (they only fill staging buffers).
 */
int cmb_wait_for_stream(cmb_ctx* ctx, void* stream);

/*
 * Trajectory ingest: like cmb_accumulate_host, but h_xyz_full holds ALL atoms,
 * [n_frames][n_total][3], and the used atoms (h_used[n_used], ascending atom indices) are gathered
 * on the fly: host threads pack chunk k+1 into pinned staging while chunk k is copied and
 * processed.  Fuses the reference's _atom_slice copy (atom_indexer.py:5-22, `xyz[:, indices]`)
 * into the streaming pipeline.  n_threads <= 0 picks a default.  Synchronous.
 */
int cmb_accumulate_host_gather(cmb_ctx* ctx, const float* h_xyz_full, int n_total, const int32_t* h_used,
                               const float* h_box, int64_t n_frames, int64_t chunk_frames, int n_threads,
                               uint32_t* d_atom_table, uint32_t* d_res_table);

/*
 * contact_trajectory._build_contacts (contact_trajectory.py:7-44): per-frame contact
 * sets as unordered 64-bit keys  (frame0+f) << 40 | lo << 20 | hi  (used-atom
 * indices for atoms, compact residue ids for residues).  d_counts[0]/[1] (device
 * uint64, zeroed by the caller) receive the number of atom / residue keys produced;
 * keys beyond the capacities are dropped (the caller retries with larger buffers).
 * Either key buffer may be NULL.  Optionally also accumulates into tables (may be NULL).
 */
int cmb_frame_contacts(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames,
                       int64_t frame0, uint64_t* d_atom_keys, int64_t atom_cap,
                       uint64_t* d_res_keys, int64_t res_cap, uint64_t* d_counts,
                       uint32_t* d_atom_table, uint32_t* d_res_table, void* stream);

/*
 * The same for a HOST trajectory (the data the reference hands to _build_contacts,
 * contact_trajectory.py:7-44, lives in host memory: trajectory.xyz): the frames stream through
 * the key kernels chunk by chunk over two streams like cmb_accumulate_host[_gather] (pinned
 * buffers are copied directly, pageable / memory-mapped ones and trajectories that still hold
 * all n_total atoms are packed into pinned staging by host threads; h_used = NULL: the
 * trajectory holds exactly the used atoms).  Every chunk APPENDS its keys -- they carry the
 * trajectory frame numbers frame0 + f -- to the caller's DEVICE buffers through the shared
 * counters d_counts (zeroed by the caller), so the result is unordered: sort it with
 * cmb_sort_keys, then cmb_keys_to_csr.  Synchronous.
 */
int cmb_frame_contacts_host(cmb_ctx* ctx, const float* h_xyz_full, int n_total, const int32_t* h_used,
                            const float* h_box, int64_t n_frames, int64_t frame0, int64_t chunk_frames,
                            int n_threads, uint64_t* d_atom_keys, int64_t atom_cap,
                            uint64_t* d_res_keys, int64_t res_cap, uint64_t* d_counts);

/*
 * Sorted keys -> the per-frame containers of ContactTrajectory (contact_trajectory.py:70-89) as
 * CSR arrays: d_ptr[n_frames + 1] (keys of frame f are [ptr[f], ptr[f+1])), d_i / d_j[n_keys] the
 * pairs, i < j, mapped through d_index_map (compact -> real index; NULL: identity).
 */
int cmb_keys_to_csr(cmb_ctx* ctx, const uint64_t* d_sorted_keys, int64_t n_keys, int64_t n_frames,
                    int64_t frame0, const int32_t* d_index_map, int64_t* d_ptr, int32_t* d_i,
                    int32_t* d_j, void* stream);

/*
 * Hashed count tables (north_star: "dense or hashed count matrix") for systems whose dense
 * [n_used][n_used] table would not stay in L2 (a 100k-atom system: 40 GB): open addressing over
 * uint64 slots, slot = (flat index + 1) << 24 | count, flat index = lo * n_used + hi as in the dense
 * layout; a zero-filled buffer is an empty table.  After cmb_set_table_mode(ctx, a, r) with a, r > 0
 * the d_atom_table / d_res_table arguments of cmb_accumulate*, cmb_frame_contacts are such slot
 * arrays of 2^a / 2^r slots (0 = dense uint32 table).  Global-memory cell path only.  Counts are
 * limited to 2^24 - 1 frames per table.  A hit that finds no free slot is PARKED in the caller's
 * overflow list (cmb_hash_overflow_buffer; atom entry = flat index + 1, residue entry =
 * 1 << 63 | (flat index + 1) << 24 | frame sequence number & 0xffffff, to be de-duplicated per
 * frame); cmb_hash_stats returns and resets the number of new keys and of parked hits since the
 * last call (it synchronises the device).  The caller grows the table (cmb_hash_rehash into a
 * larger zero-filled buffer) and adds the parked hits with cmb_hash_add(keys + 1, counts); parked
 * hits beyond the list capacity are lost, which the caller must treat as an error.
 */
int cmb_set_table_mode(cmb_ctx* ctx, int atom_log2, int res_log2);
int cmb_hash_overflow_buffer(cmb_ctx* ctx, uint64_t* d_list, int64_t cap);
int cmb_hash_stats(cmb_ctx* ctx, int64_t* new_atom_keys, int64_t* new_res_keys, int64_t* overflow);
int cmb_hash_add(cmb_ctx* ctx, uint64_t* d_slots, int log2_slots, const uint64_t* d_key1,
                 const uint64_t* d_count, int64_t n, int which, void* stream);
int cmb_export_hashed(cmb_ctx* ctx, const uint64_t* d_slots, int log2_slots, int64_t cap,
                      int64_t* d_idx, uint32_t* d_cnt, uint64_t* d_n, void* stream);
int cmb_hash_rehash(cmb_ctx* ctx, const uint64_t* d_src, int src_log2, uint64_t* d_dst, int dst_log2,
                    void* stream);

/* In-place ascending radix sort of 64-bit keys (frame-major, then lo, then hi). */
int cmb_sort_keys(cmb_ctx* ctx, uint64_t* d_keys, int64_t n, void* stream);

/*
 * Sparse export of a count table: flat indices and counts of the non-zero entries
 * (unordered).  *d_n (device uint64, zeroed by the caller) receives the number found;
 * entries beyond cap are dropped.  Replaces the Counter objects of
 * contact_map.py:724-744 as the result container.
 */
int cmb_export_nonzero(cmb_ctx* ctx, const uint32_t* d_table, int64_t n_elems, int64_t cap,
                       int64_t* d_idx, uint32_t* d_cnt, uint64_t* d_n, void* stream);

/*
 * Sorted sparse export (the form the Python layer consumes): d_packed[cap] receives
 * (flat index << 24 | count) of the non-zero entries in ascending index order (dense tables: an
 * ordered two-pass scan, no sort; hashed tables: unordered scan + radix sort on the index bits,
 * unused tail of d_packed = 0xff..ff); d_n[0] = entries found (those beyond cap are dropped: retry
 * with a larger buffer),
 * d_n[1] != 0 if a count did not fit 24 bits (use cmb_export_nonzero then).  log2_slots > 0:
 * d_table is a hashed table of 2^log2_slots slots, else a dense uint32 table of n_elems entries.
 */
int cmb_export_sorted(cmb_ctx* ctx, const void* d_table, int64_t n_elems, int log2_slots, int64_t cap,
                      uint64_t* d_packed, uint64_t* d_n, void* stream);

/*
 * Split the packed words of cmb_export_sorted into the two arrays a host consumer wants (the keys and
 * values of the reference's Counter, contact_map.py:739-740): d_idx[k] = word >> 24 (+ idx_offset),
 * d_cnt[k] = word & 0xffffff for k < min(d_n[0], cap).  The entry count is read on the device, so the
 * call can be queued right behind cmb_export_sorted without a host synchronisation in between.
 */
int cmb_unpack_sorted(cmb_ctx* ctx, const uint64_t* d_packed, const uint64_t* d_n, int64_t cap,
                      int64_t idx_offset, int64_t* d_idx, uint32_t* d_cnt, void* stream);

/*
 * The same for the export of ONE scan over an allocation that holds two dense tables (the atom table
 * of n_cols0 columns at flat index 0, the residue table of n_cols1 columns at flat index off1): every
 * entry is split into its (row, column) pair on the device -- d_lo[k], d_hi[k] (compact indices, int32),
 * d_cnt[k] -- and d_split[0] receives the number of entries that belong to the first table (they come
 * first: the words are sorted).  Replaces the per-key Python loop that fills the reference's Counter
 * (contact_map.py:739-740) and a host-side divmod over millions of entries.
 */
int cmb_unpack_pairs(cmb_ctx* ctx, const uint64_t* d_packed, const uint64_t* d_n, int64_t cap,
                     int64_t n_cols0, int64_t off1, int64_t n_cols1, int32_t* d_lo, int32_t* d_hi,
                     uint32_t* d_cnt, uint64_t* d_split, void* stream);

/*
 * north_star: "any pair within 1 ulp of the cutoff enumerated and reported".
 * Records {frame, lo, hi, d2} (int32,int32,int32,float32 = 16 bytes) of eligible
 * pairs whose d2 is within `ulps` float steps of cutoff^2.
 */
int cmb_boundary_pairs(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames,
                       int64_t frame0, int ulps, void* d_records, int64_t cap, uint64_t* d_n,
                       void* stream);

/*
 * Audit of the cell kernels' screening shortcut: the kernels call a pair a contact without running
 * MDTraj's arithmetic when its screening distance is below a per-frame threshold c2_lo, and drop it
 * when it is above c2_hi (md.compute_neighborlist, contact_map.py:575, is then never evaluated for
 * that pair).  This entry runs the reference arithmetic on EVERY stencil candidate that is not
 * residue-excluded and compares: d_out[0] (device uint64) = pairs audited, d_out[1] = pairs the
 * screen alone would have decided differently -- which must be 0.  Same frame arguments as
 * cmb_accumulate.  The pair-list path has its own audit (cmb_set_option "verlet_audit").
 */
int cmb_screen_audit(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames,
                     uint64_t* d_out, void* stream);

/* Executed candidate-pair tests of the last accumulate / frame_contacts call when
 * enabled with cmb_set_option(ctx, "count_tests", 1) (bench metric; synchronises).
 * Other options: "force_path" (-1 auto, 0 fused kernel, 1 global-memory cell path), "max_cells",
 * "pair_tasks", "stage_pageable", "large_batch" (global-memory cell path: 0 = one frame at a time
 * with frame stamps, 1 = many frames per launch with per-frame residue bitmaps when the system
 * is small enough for that to pay (default), 2 = batched whenever the bitmaps fit). */
int cmb_set_option(cmb_ctx* ctx, const char* name, int64_t value);
int cmb_get_stat(cmb_ctx* ctx, const char* name, int64_t* value);

/*
 * AtomContactConcurrence / ResidueContactConcurrence (concurrence.py:100-162):
 * d_out[f][g] = 1 iff min over the atom pairs of group g of the compute_distances() distance
 * (float32, minimum image as in min_dist) is < cutoff (float32 compare, strict) in frame f.
 * d_atoms[n_sel]: the atoms the pairs touch; d_pairs[P][2]: indices INTO d_atoms;
 * d_group_ptr[n_groups + 1]: CSR offsets of the groups in d_pairs (an atom contact is a group of
 * one pair, a residue contact the product of the two residues' selected atoms).
 */
int cmb_group_contacts(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames, int n_atoms,
                       const int32_t* d_atoms, int n_sel, const int32_t* d_pairs,
                       const int64_t* d_group_ptr, int n_groups, int orthogonal, float cutoff,
                       uint8_t* d_out, void* stream);

/*
 * _atom_slice (atom_indexer.py:5-22): out[f][i][:] = xyz[f][used[i]][:] on device.
 */
int cmb_gather_atoms(cmb_ctx* ctx, const float* d_xyz_full, int64_t n_frames, int n_total,
                     const int32_t* d_used, int n_used, float* d_out, void* stream);

/*
 * MinimumDistanceCounter._compute_from_atom_pairs (min_dist.py:142-168) for the
 * query-major product query x haystack (min_dist.py:138): per frame the minimum
 * float32 distance (compute_distances arithmetic) and the FIRST pair index
 * (q_pos * n_h + h_pos) attaining it.  d_xyz is the FULL trajectory
 * [n_frames][n_atoms][3]; orthogonal = np.allclose(unitcell_angles, 90).
 */
int cmb_min_dist(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames,
                 int n_atoms, const int32_t* d_query, int n_q, const int32_t* d_haystack, int n_h,
                 int orthogonal, float* d_min, int64_t* d_arg, void* stream);

/*
 * MinimumDistanceCounter._compute_from_atom_pairs(trajectory, atom_pairs) (min_dist.py:142-168)
 * for an ARBITRARY pair list d_pairs[n_pairs][2] (int32 atom indices, device): per frame
 * D.min(axis=1) and D.argmin(axis=1) of D = md.compute_distances(trajectory, atom_pairs)
 * (min_dist.py:165-167) without materialising D; ties -> lowest position in the list.
 */
int cmb_min_dist_pairs(cmb_ctx* ctx, const float* d_xyz, const float* d_box, int64_t n_frames,
                       int n_atoms, const int32_t* d_pairs, int64_t n_pairs, int orthogonal,
                       float* d_min, int64_t* d_arg, void* stream);

/*
 * NearestAtoms._calculate_nearest (min_dist.py:47-70) for one frame: for every atom
 * the nearest allowed neighbour within the cutoff (neighbour-list arithmetic for the
 * cutoff test, compute_distances arithmetic for the value; ties -> lower index).
 * excl_mode 0: only the atom itself is excluded (excluded == {});
 *           1: atoms of the same residue (excluded is None), d_atom_res[n_atoms];
 *           2: CSR lists d_excl_ptr[n_atoms+1] / d_excl_idx.
 * d_nearest[i] = -1 where no allowed neighbour exists.
 */
int cmb_nearest(cmb_ctx* ctx, const float* d_xyz_frame, const float* d_box9, int n_atoms,
                double cutoff, int orthogonal, int excl_mode, const int32_t* d_atom_res,
                const int64_t* d_excl_ptr, const int32_t* d_excl_idx, int32_t* d_nearest,
                float* d_dist, void* stream);

/*
 * Bit-reproducible synthetic trajectory generator, device side:
 * x = base + A * (2u - 1) + image shift per residue, u = splitmix64 hash.  Matches
 * contact_map_b200/synth.py bit for bit.
 */
int cmb_synth_frames(cmb_ctx* ctx, const float* d_base, const int32_t* d_atom_group, int n_atoms,
                     const float* d_box9, uint64_t seed, float amplitude, int image_shifts,
                     int64_t frame0, int64_t n_frames, float* d_out, void* stream);

/*
 * Device memory for hosts that bring no CUDA allocator of their own (the Python layer uses torch
 * tensors instead): zero-filled allocation, free, asynchronous zero fill, synchronous copies.
 * A zero-filled buffer of cmb_table_sizes() elements is an empty count table.
 */
int cmb_device_alloc(cmb_ctx* ctx, uint64_t n_bytes, void** d_ptr);
int cmb_device_free(cmb_ctx* ctx, void* d_ptr);
int cmb_device_zero(cmb_ctx* ctx, void* d_ptr, uint64_t n_bytes, void* stream);
int cmb_copy_to_host(cmb_ctx* ctx, void* h_dst, const void* d_src, uint64_t n_bytes, void* stream);
int cmb_copy_to_device(cmb_ctx* ctx, void* d_dst, const void* h_src, uint64_t n_bytes, void* stream);

/*
 * reduce_all_results (frequency_task.py:125-141; orchestrated by dask_run, dask_runner.py:11-37)
 * across the GPUs of a job, one process per GPU: the ranks' dense uint32 count tables -- the layout
 * is identical on every rank -- are summed in place with ONE NCCL collective over NVLink/NVSwitch.
 * root < 0: all-reduce (every rank ends up with the sum); root >= 0: reduce to that rank only,
 * which is what the reference's client-side reduce delivers.  `comm` is an ncclComm_t.  NCCL is
 * bound at run time (dlopen "libnccl.so.2", or the path in $CMB200_NCCL), so the library loads
 * without it; cmb_nccl_unique_id / cmb_nccl_comm_create / cmb_nccl_comm_destroy wrap
 * ncclGetUniqueId / ncclCommInitRank / ncclCommDestroy for hosts without NCCL bindings (rank 0
 * creates the id and hands its 128 bytes to the other ranks by whatever means the host has).
 * Hashed tables have rank-dependent layouts: export them and merge the sparse lists instead.
 */
typedef struct { char internal[128]; } cmb_nccl_id;
int cmb_nccl_unique_id(cmb_ctx* ctx, cmb_nccl_id* out);
int cmb_nccl_comm_create(cmb_ctx* ctx, int n_ranks, const cmb_nccl_id* id, int rank, void** comm);
int cmb_nccl_comm_destroy(cmb_ctx* ctx, void* comm);
int cmb_reduce(cmb_ctx* ctx, void* comm, uint32_t* d_table, int64_t n_elems, int root, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMB200_H */
