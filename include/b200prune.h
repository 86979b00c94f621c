/*
 * b200prune.h — C-ABI of the B200-native pruning hot path.
 *
 * Drop-in boundary for EIDOSLAB/pruning-for-vision-representation's pruning path.  The
 * reference has no FFI of its own (pure Python over torch ops); each entry point below
 * names the reference lines whose arithmetic it replaces.  A reference maintainer binds
 * this header with ctypes; `pruning_for_vision_representation_b200`
 * is that binding plus a mirror of the reference's Python functions.
 *
 * Conventions
 *   - every pointer named d_* is DEVICE memory owned by the caller (PyTorch); h_* is host.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *   - all entry points are asynchronous w.r.t. the host unless they say "synchronises".
 *   - return value: 0 on success, negative B200P_E* on failure; b200p_last_error() gives
 *     a thread-local message.  There is no CPU fallback: without a CUDA device every
 *     compute entry point fails with B200P_ECUDA.
 *
 * Data model ("plan")
 *   The prunable parameter set is a list of T segments (one per nn.Conv2d / nn.Linear
 *   weight, in model.named_modules() order — train.py:263-265, 333-336).  The virtual
 *   flat index space is their concatenation (the reference's torch.cat, train.py:294 /
 *   parameters_to_vector, torch/nn/utils/prune.py:1114).  Each segment is cut into
 *   chunks of B200P_CHUNK elements; a chunk never spans two segments.  The bit-packed
 *   mask has one 32-bit word per 32 elements, chunk-major: chunk c owns words
 *   [c*128, c*128+128); bit (e & 31) of word (e >> 5) inside the chunk is element e.
 *   Bits past a segment's last element are always 0.  Bit = 1 means "kept".
 */
#ifndef B200PRUNE_H_
#define B200PRUNE_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200P_CHUNK            4096
#define B200P_WORDS_PER_CHUNK  128

#define B200P_OK        0
#define B200P_EINVAL   -1   /* bad argument (shape / alignment / range)        */
#define B200P_ECUDA    -2   /* CUDA runtime error or no device                   */
#define B200P_ESTATE   -3   /* call sequence error (e.g. slot not bound)          */
#define B200P_ENOMEM   -4

/* pointer-table slots of a plan: one device pointer per segment */
#define B200P_SLOT_W        0   /* fp32 weights (weight_orig)                              */
#define B200P_SLOT_G        1   /* fp32 gradients                                          */
#define B200P_SLOT_SCORE    2   /* fp32 accumulated SNIP scores                            */
#define B200P_SLOT_BUF      3   /* fp32 SGD momentum buffers                               */
#define B200P_SLOT_WEFF     4   /* fp32 effective (masked) weight = module.weight          */
#define B200P_SLOT_MASKF    5   /* fp32 0/1 masks = module.weight_mask (checkpoint compat) */
#define B200P_SLOT_WEFF16   6   /* bf16 effective weight for autocast forward              */
#define B200P_SLOT_EMA      7   /* fp32 EMA of weight_orig (utils.py:159-170)               */
#define B200P_NUM_SLOTS     8

/* emit / select modes */
#define B200P_MODE_SNIP_STRICT  0  /* keep = score > thr (every tie pruned): train.py:316          */
#define B200P_MODE_EXACT_K      1  /* prune exactly k alive entries, ties lowest flat index first: */
                                   /* prune.L1Unstructured, torch/nn/utils/prune.py:520-540        */

/* key source for select / emit */
#define B200P_KEY_ABS_W   0   /* key = |w|  (magnitude pruning; |.| is fused into the read) */
#define B200P_KEY_SCORE   1   /* key = SCORE slot (SNIP)                                    */

typedef struct b200p_plan b200p_plan;

/* result block written by the select stage (device memory inside the workspace; copy it
 * out with b200p_select_result).  Mirrors the Python float `threshold` of train.py:307
 * plus the tie bookkeeping of SURVEY §8(c). */
typedef struct b200p_select_result_t {
    uint64_t k;          /* requested rank (1-based k-th smallest among alive keys)          */
    uint64_t n_valid;    /* alive keys considered                                            */
    uint64_t n_less;     /* alive keys strictly below the threshold                          */
    uint64_t n_equal;    /* alive keys equal to the threshold                                */
    uint64_t quota;      /* EXACT_K: how many of the tied keys are pruned (= k - n_less)     */
    uint64_t n_kept;     /* kept bits after the last emit                                    */
    float    threshold;  /* the k-th smallest key as fp32 (NaN if the rank falls on a NaN)   */
    uint32_t thr_key;    /* its 31-bit integer key                                           */
    uint32_t passes_full;/* how many full-data passes the select needed (2 or 3)             */
    uint32_t collected;  /* candidates gathered in pass 2 (0 if histogram mode)              */
    uint32_t miss;       /* sampled select only: 1 = the bracket missed rank k.  One GPU: the exact */
                         /* select ran inside the same launch, the result is valid.  Sharded        */
                         /* (b200p_sharded_mask_build): the result is NOT valid, rerun staged.       */
    uint32_t reserved_;
} b200p_select_result_t;

const char* b200p_last_error(void);
int  b200p_version(void);
/* number of visible CUDA devices (0 if none / driver missing); never fails */
int  b200p_device_count(void);

/* ---- plan ------------------------------------------------------------------------- */
/* Build the chunk tables for T segments with the given element counts (host array).
 * Allocates the small device tables and a workspace (histograms, select state, candidate
 * buffer of `cand_capacity` entries; 0 = default max(1<<20, N/16)).  Synchronises. */
int  b200p_plan_create(int device, int n_segments, const int64_t* h_numel,
                       int64_t cand_capacity, b200p_plan** out);
int  b200p_plan_destroy(b200p_plan* plan);
int64_t b200p_plan_total(const b200p_plan* plan);        /* N = sum numel               */
int64_t b200p_plan_num_chunks(const b200p_plan* plan);
int64_t b200p_plan_mask_words(const b200p_plan* plan);   /* = num_chunks * 128            */
int64_t b200p_plan_seg_chunk_start(const b200p_plan* plan, int seg); /* first chunk of seg */
/* flat (segment-concatenated, unpadded) index of the first element of chunk c; c = num_chunks
 * gives N.  A chunk range [c0,c1) is the contiguous flat range [flat(c0), flat(c1)). */
int64_t b200p_plan_chunk_flat_start(const b200p_plan* plan, int64_t chunk);
/* Upload one device pointer per segment for `slot`.  16-byte aligned pointers take the
 * 128-bit vector path; anything else is accepted and runs the scalar path. */
int  b200p_plan_bind(b200p_plan* plan, int slot, const void* const* h_ptrs, void* stream);
/* Reusable pointer table: the same per-segment pointers expanded once into a device-resident
 * per-chunk table.  Binding it to a slot is a host-side swap (no launch, no copy), which is what
 * a loop over mini-batches wants (one table per gradient set).  `slot` fixes the element size.
 * The table must outlive every launch that uses it. */
typedef struct b200p_ptrtable b200p_ptrtable;
int  b200p_ptrtable_create(b200p_plan* plan, int slot, const void* const* h_ptrs, void* stream,
                           b200p_ptrtable** out);
int  b200p_ptrtable_destroy(b200p_ptrtable* table);
/* Re-point n existing tables (of one plan, same slot) at new per-segment pointers — the fresh gradient tensors autograd
 * hands over for the next mask build (train.py:258-269) — in ONE small launch: the pointers travel as kernel arguments
 * (up to 1024 per launch), nothing is allocated, nothing synchronises.  h_ptrs[i] is the pointer list of tables[i]. */
int  b200p_ptrtables_update(b200p_ptrtable* const* tables, const void* const* const* h_ptrs, int n_tables, int slot, void* stream);
int  b200p_plan_bind_table(b200p_plan* plan, int slot, const b200p_ptrtable* table);
/* options */
#define B200P_OPT_SELECT_IMPL   1
#define B200P_SELECT_SAMPLED    0   /* default: 1/64 sample -> bracket -> one full pass -> candidates; exact fallback on a miss */
#define B200P_SELECT_EXACT      1   /* 3-pass MSD radix select over the full data */
#define B200P_OPT_TIME_SWEEP    2   /* value 1: record CUDA events around every fused score+sweep kernel of
                                       b200p_snip_mask_build / b200p_snip_score_select (measurement only) */
#define B200P_OPT_REUSE_SAMPLE  3   /* value 1: selects over UNCHANGED keys (same key source, old mask, chunk range; e.g. a sparsity
                                       sweep over fixed weights, BASELINE config 5) derive their bracket from the sample histogram the
                                       first one cached instead of sampling (and, sharded, all-reducing) again.  The caller vouches
                                       for "unchanged": weights updated in place must clear the option or re-bind the slot. */
#define B200P_OPT_COOP_GRID     4   /* cap on the grid of this plan's cooperative launches (0 = one CTA per SM).  Several plans that share
                                       ONE device and wait on each other (virtual ranks in tests) must fit on it together. */
int  b200p_plan_set_option(b200p_plan* plan, int option, int64_t value);
/* Mean duration (ms) and number of the fused score+sweep launches timed since the option was set or this was last
 * called; synchronises on the last recorded event.  *out_launches may be 0 (then *out_ms = 0). */
int  b200p_plan_kernel_time_ms(b200p_plan* plan, double* out_ms, int64_t* out_launches);
/* device address of the 4096-bin uint64 histogram and of the select state, so that the
 * host side can run NCCL collectives on them between stages */
void* b200p_plan_hist_ptr(b200p_plan* plan);
void* b200p_plan_state_ptr(b200p_plan* plan);

/* ---- K1: importance scores (train.py:258-261, 282-291) ----------------------------- */
/* SCORE[t] (=|+=) |W[t] * G[t]|   (accumulate=0 assigns, 1 adds) over chunks [c0,c1) */
int  b200p_score_accumulate(b200p_plan* plan, int accumulate,
                            int64_t chunk_begin, int64_t chunk_end, void* stream);

/* The same over n_sets gradient sets in ONE pass (groups of 8 per launch): SCORE (=|+=) sum_b |W*G_b|,
 * added in table order — bit-identical to n_sets calls of b200p_score_accumulate, but W is read once and
 * SCORE touched once: 4*(n_sets+2) B/param instead of 16*n_sets.  g_tables: b200p_ptrtable_create(slot G). */
int  b200p_score_accumulate_multi(b200p_plan* plan, const b200p_ptrtable* const* g_tables, int n_sets, int accumulate,
                                  int64_t chunk_begin, int64_t chunk_end, void* stream);

/* Multi-GPU score exchange: d_dst[i] = ((d_src[i] + d_src[stride+i]) + ...) over
 * n_parts partial score slices received from the ranks, summed in rank (= mini-batch) order. */
int  b200p_sum_parts(int device, float* d_dst, const float* d_src, int n_parts, int64_t part_stride,
                     int64_t n, void* stream);

/* ---- K2: global k-th smallest (train.py:299-307; prune.py:526-536) ----------------- */
/* Whole select on one GPU: radix passes + scans, no host sync.  `d_old_mask` (nullable)
 * restricts the key set to alive entries (iterative magnitude pruning, prune.py:368-370).
 * k is the 1-based rank among alive keys, 1 <= k <= n_alive. */
int  b200p_select_kth(b200p_plan* plan, int key_source, const uint32_t* d_old_mask,
                      uint64_t k, int mode, void* stream);
/* staged variant for parameter-sharded multi-GPU select: the caller runs
 *   begin -> [hist(pass) -> allreduce(hist) -> scan(pass)] for pass = 0,1,2 -> finish
 * over its own chunk range; hist/state pointers come from b200p_plan_hist_ptr/state_ptr. */
int  b200p_select_begin(b200p_plan* plan, uint64_t k, int mode, int allow_collect, void* stream);
int  b200p_select_hist(b200p_plan* plan, int pass, int key_source, const uint32_t* d_old_mask,
                       int64_t chunk_begin, int64_t chunk_end, void* stream);
int  b200p_select_scan(b200p_plan* plan, int pass, void* stream);
/* EXACT_K tie resolution over chunks [c0,c1): `tie_offset` = number of tied keys that
 * live in lower-numbered chunks owned by other ranks (0 on one GPU). */
int  b200p_select_ties(b200p_plan* plan, int key_source, const uint32_t* d_old_mask,
                       int64_t chunk_begin, int64_t chunk_end, uint64_t tie_offset, void* stream);
/* The same in two stages for the parameter-sharded multi-GPU select: `count` writes the number
 * of tied keys inside [c0,c1) to caller-owned device memory (8 bytes) so that the host side can
 * all-gather it; `scan` takes the gathered per-rank table and adds the counts of the n_before
 * lower ranks on the device (no host round trip). */
int  b200p_select_ties_count(b200p_plan* plan, int key_source, const uint32_t* d_old_mask,
                             int64_t chunk_begin, int64_t chunk_end, uint64_t* d_local_count, void* stream);
int  b200p_select_ties_scan(b200p_plan* plan, int64_t chunk_begin, int64_t chunk_end,
                            const uint64_t* d_counts, int n_before, void* stream);
/* copy the result block to the host.  Synchronises the stream. */
int  b200p_select_result(b200p_plan* plan, b200p_select_result_t* h_out, void* stream);

/* ---- K3: mask emit (train.py:311-317; prune.py:538, 1149-1161) ---------------------- */
/* new_mask = old_mask & keep(key, threshold, mode).  Optional fused outputs when the
 * corresponding slots are bound and requested: MASKF (fp32 0/1), WEFF (= mask ? W : 0).
 * force: 0 = use the selected threshold, 1 = keep everything alive, 2 = prune everything,
 *        3 = SNIP_STRICT against `forced_threshold` (train.py:300-303: +inf / -1). */
#define B200P_EMIT_MASKF   1
#define B200P_EMIT_WEFF    2
int  b200p_emit_masks(b200p_plan* plan, int key_source, int mode, int force,
                      float forced_threshold, const uint32_t* d_old_mask, uint32_t* d_new_mask,
                      int outputs, int64_t chunk_begin, int64_t chunk_end, void* stream);

/* Select + emit in one call (k-th smallest of the alive keys, then the packed mask).  Same result as
 * b200p_select_kth followed by b200p_emit_masks; the select's sweep writes a provisional mask directly into
 * d_new_mask and the emit only patches the ~2 % of keys near the threshold instead of re-reading all keys. */
int  b200p_mask_build(b200p_plan* plan, int key_source, const uint32_t* d_old_mask, uint64_t k, int mode,
                      uint32_t* d_new_mask, void* stream);

/* SNIP score + select + emit in one call (train.py:282-317 with the multi-batch score of SURVEY 8c):
 * SCORE = sum over the n_sets gradient tables of |W * G_b| (written to the SCORE slot, bit-identical to
 * b200p_score_accumulate_multi), threshold = k-th smallest score, mask = score > threshold (strict).  Same
 * result as b200p_score_accumulate_multi followed by b200p_mask_build(KEY_SCORE, NULL, k, SNIP_STRICT), but the
 * pass that writes the scores also classifies them against a bracket obtained from a 1/64 sample, so the scores
 * are never read back (a bracket miss runs the exact select over the SCORE slot).  W and SCORE must be bound. */
int  b200p_snip_mask_build(b200p_plan* plan, const b200p_ptrtable* const* g_tables, int n_sets, uint64_t k,
                           uint32_t* d_new_mask, void* stream);
/* The same for gradient tensors that are NEW since the tables were filled — the normal case: every backward pass
 * (train.py:270-280) allocates its gradients.  h_ptrs[i][t] = device pointer of segment t of gradient set i
 * (as for b200p_ptrtables_update).  The sample kernel writes the per-chunk tables on its way, so the build is
 * one launch shorter than b200p_ptrtables_update + b200p_snip_mask_build; it falls back to that pair when
 * n_sets > 8, n_sets * n_segments > 1024, or the fused path is unavailable. */
int  b200p_snip_mask_build_refresh(b200p_plan* plan, b200p_ptrtable* const* g_tables, const void* const* const* h_ptrs,
                                   int n_sets, uint64_t k, uint32_t* d_new_mask, void* stream);
/* The score + select half of it (= b200p_score_accumulate_multi + b200p_select_kth(KEY_SCORE, NULL, k, SNIP_STRICT)):
 * the caller issues its own b200p_emit_masks afterwards (e.g. with an old mask to AND with, or fp32 mask outputs).
 * d_prov_target: nullable, where the provisional mask of the sweep goes (the d_new_mask of a following plain emit). */
int  b200p_snip_score_select(b200p_plan* plan, const b200p_ptrtable* const* g_tables, int n_sets, uint64_t k,
                             uint32_t* d_prov_target, void* stream);

/* ---- parameter-sharded mask build over peer memory (no reference counterpart: under DDP the reference's
 * ranks prune independently and end up with different masks, train.py:604-607, 622-628) ------------------------------
 * One process per GPU.  Every rank creates a comm of the same geometry; the peer-visible WINDOW (flags, histogram /
 * gather slots, the full packed mask, a score area of `score_cap` floats per source rank) is exchanged as a CUDA IPC
 * handle (64 bytes, e.g. through torch.distributed.all_gather) or, for plans that live on one device, as plain
 * pointers.  The kernels then talk through the windows directly: the last CTA of the select's sample and sweep kernels
 * all-reduces its histogram by pushing it into every peer's window, the finish kernel all-gathers 4 KB, the emit's
 * result is pushed as packed words — no NCCL call, no host round trip inside a build.  At most 8 ranks (one node). */
typedef struct b200p_comm b200p_comm;
int  b200p_comm_create(int device, int rank, int world, int64_t mask_words, int64_t score_cap, b200p_comm** out);
int  b200p_comm_destroy(b200p_comm* comm);
/* Measurement aid: globaltimer (ns) stamps of the last in-kernel collective per channel and sequence parity
 * (comm.cuh: comm_signal_and_wait), u64 [4][2][8].  Synchronises the device. */
int  b200p_comm_trace(b200p_comm* c, uint64_t* h_out64);
int64_t b200p_comm_window_bytes(const b200p_comm* comm);
void* b200p_comm_window(const b200p_comm* comm);                 /* device address of the own window            */
void* b200p_comm_mask_ptr(const b200p_comm* comm);               /* the full packed mask inside the own window  */
void* b200p_comm_score_ptr(const b200p_comm* comm);              /* score area: world parts of score_cap floats */
int64_t b200p_comm_score_cap(const b200p_comm* comm);
int  b200p_comm_ipc_handle(b200p_comm* comm, void* h_out64);     /* cudaIpcMemHandle_t of the own window         */
int  b200p_comm_connect_ipc(b200p_comm* comm, const void* h_handles /* world x 64 bytes, rank order */);
int  b200p_comm_connect_local(b200p_comm* comm, void* const* h_peer_windows /* world device pointers */);
/* 0 = fine, ch + 1 = a wait on channel ch ran out (a peer never arrived; bounded spin, ~2 s).  Synchronises. */
int  b200p_comm_error(b200p_comm* comm, void* stream);
int  b200p_comm_barrier(b200p_comm* comm, void* stream);
/* Every rank pushes the mask words of its chunk range (already in its window) into every peer's window; when the kernel
 * ends on a rank, all slices have arrived there: b200p_comm_mask_ptr() holds the full mask. */
int  b200p_comm_mask_allgather(b200p_comm* comm, b200p_plan* plan, int64_t chunk_begin, int64_t chunk_end, void* stream);
/* SCORE-slot pointer table whose entries point into the OWNING ranks' score areas (h_bounds: world + 1 chunk bounds):
 * with it bound, b200p_score_accumulate(_multi) writes each chunk's partial scores straight into the owner's window —
 * the local score pass and the score exchange are one kernel.  Follow with b200p_comm_barrier and b200p_sum_parts
 * (parts added in rank = mini-batch order). */
int  b200p_comm_score_push_table(b200p_comm* comm, b200p_plan* plan, const int64_t* h_bounds, void* stream, b200p_ptrtable** out);
/* k-th smallest over ALL ranks' chunk ranges + packed mask of the whole set in every window: sample -> sweep of the own
 * range -> exact key -> ties -> emit -> mask all-gather, see csrc/select.cu.  Same results as b200p_mask_build on one GPU
 * (bit-identical masks, threshold, tie policy).  Check b200p_select_result().miss before using the mask. */
/* `stages`: 0 = the whole sequence; a bit set runs single stages (in this order), which is what lets several ranks that
 * live in ONE process be interleaved stage by stage on their own streams (tests with virtual ranks). */
#define B200P_SHARD_SAMPLE  1
#define B200P_SHARD_SWEEP   2
#define B200P_SHARD_FINISH  4
#define B200P_SHARD_TIES    8
#define B200P_SHARD_EMIT   16
#define B200P_SHARD_PUSH   32
#define B200P_SHARD_ALL    63
int  b200p_sharded_mask_build(b200p_plan* plan, b200p_comm* comm, int key_source, const uint32_t* d_old_mask, uint64_t k, int mode,
                              int64_t chunk_begin, int64_t chunk_end, int stages, void* stream);

/* ---- K5: sparsity (train.py:347-369) ------------------------------------------------ */
/* d_out[0] = # elements with (mask bit == 0 or W == 0)  (d_mask nullable -> counts W == 0),
 * d_out[1] = # mask bits set.  d_out is 2 x uint64 device memory, zeroed by the call. */
int  b200p_count_zeros(b200p_plan* plan, const uint32_t* d_mask, uint64_t* d_out,
                       int use_weights, void* stream);
/* fp32 0/1 masks -> packed (checkpoint load) and packed -> fp32 MASKF slot */
int  b200p_mask_pack_from_f32(b200p_plan* plan, uint32_t* d_mask, void* stream);
int  b200p_mask_unpack_to_f32(b200p_plan* plan, const uint32_t* d_mask, void* stream);
/* WEFF = mask ? W : 0 (the forward pre-hook of prune.py:71-74), optionally bf16 too */
int  b200p_apply_mask(b200p_plan* plan, const uint32_t* d_mask, int outputs, void* stream);
/* G = mask ? G : 0 (the MulBackward of the reparametrisation) */
int  b200p_mask_grads(b200p_plan* plan, const uint32_t* d_mask, void* stream);

/* ---- K4: fused masked SGD step (train.py:54-67; torch/optim/sgd.py:343-380) -------- */
#define B200P_SGD_NESTEROV    1
#define B200P_SGD_FIRST_STEP  2   /* momentum buffer is initialised to the gradient */
#define B200P_SGD_EMIT_WEFF   4   /* also write WEFF   = mask ? w_new : 0           */
#define B200P_SGD_EMIT_WEFF16 8   /* also write WEFF16 = bf16(mask ? w_new : 0)     */
int  b200p_masked_sgd_step(b200p_plan* plan, const uint32_t* d_mask, float lr, float momentum,
                           float dampening, float weight_decay, int flags, void* stream);
/* The same with a device-side control block d_ctl (2 floats, nullable): every gradient is multiplied by d_ctl[0] before
 * use — the loss-scale removal of GradScaler.unscale_, the coefficient of clip_grad_norm_ and the 1/world of a gradient
 * all-reduce folded into one factor (train.py:55-66) — and d_ctl[1] != 0 skips the whole step (a non-finite gradient
 * was found: GradScaler.step).  Both are read on the device, so nothing between backward and step needs the host. */
int  b200p_masked_sgd_step_ctl(b200p_plan* plan, const uint32_t* d_mask, float lr, float momentum, float dampening,
                               float weight_decay, int flags, const float* d_ctl, void* stream);
/* d_out2[0] = sum of g^2 over the KEPT entries of the G slot (= ||weight_orig.grad||^2 of the reference's masked
 * gradients, the prunable tensors' share of clip_grad_norm_'s total norm), d_out2[1] = number of non-finite gradient
 * entries (kept or pruned: inf * 0 = NaN reaches the reference's inf check too).  d_out2: 2 doubles, zeroed by the call. */
int  b200p_grad_stats(b200p_plan* plan, const uint32_t* d_mask, double* d_out2, void* stream);
/* EMA slot = copy ? W : decay * EMA + (1 - decay) * W   (ExponentialMovingAverage over weight_orig, utils.py:159-170;
 * copy = the n_averaged == 0 case of torch.optim.swa_utils.AveragedModel.update_parameters) */
int  b200p_ema_update(b200p_plan* plan, float decay, int copy, void* stream);

/* ---- LOST (object_discovery.py:23-134) ---------------------------------------------- */
/* Batched LOST over B images that share d.  Image b has n_b = dims[2b]*dims[2b+1] patch
 * keys stored row-major at d_feats + feat_offset[b] (elements), row stride `row_stride`
 * elements (so a qkv buffer can be read in place, main_lost_original.py:251-263).
 * Outputs (device): degree int32 [sum n_b] (the call zeroes d_degree[min out_offset, max out_offset
 * + n_b) first), seed int32 [B], box float [B,4] (xmin,ymin,xmax,ymax in pixels, exact integers when
 * the scales are integers, object_discovery.py:120-128), status int32 [B]
 * (0 ok, 1 = "The seed is in the background component", object_discovery.py:110-111; 2 = internal: the finish
 * kernel gave up waiting for the Gram kernel, never expected).
 * d_A (nullable): fp32 Gram matrices, image b at d_A + a_offset[b], n_b x n_b row-major.  With d_A == NULL and a pair
 * Gram (TC2 / TC2D) NO Gram matrix is materialised anywhere: the Gram epilogue only counts, and
 * A[seed, potentials] and M = K (sum of the similar keys) come from the keys (object_discovery.py:61-62 as two skinny
 * mat-vecs; same signs wherever an entry is decidable in fp32).
 * h_meta is a host array of B records; it is consumed before the call returns.  At most 4096
 * patches per image and k_patches <= 1024.  Workspace: b200p_lost_workspace_bytes(). */
typedef struct b200p_lost_image_t {
    int64_t feat_offset;   /* elements from d_feats                      */
    int64_t a_offset;      /* elements from d_A (ignored if d_A == NULL) */
    int64_t out_offset;    /* elements from d_degree                     */
    int32_t dim0, dim1;    /* dims = [w_featmap, h_featmap] (object_discovery.py:97) */
    int32_t img_h, img_w;  /* init_image_size[1:] (object_discovery.py:66,126-128)   */
    float   scale0, scale1;/* scales (object_discovery.py:120-121)                   */
} b200p_lost_image_t;

#define B200P_LOST_GRAM_FFMA   0   /* fp32 CUDA-core Gram                                   */
#define B200P_LOST_GRAM_TC     1   /* TMA + tcgen05 3xTF32 error-compensated Gram, one CTA per 128x128 tile */
#define B200P_LOST_GRAM_TC2    2   /* the same on CTA pairs: tcgen05.mma.cta_group::2, 256x256 tiles, features pre-split into hi/lo arrays */
#define B200P_LOST_GRAM_TC2D   3   /* default: CTA pairs reading the caller's features in place (TMA on d_feats, hi/lo tiles derived in
                                      shared memory); falls back to TC2 when the layout is not TMA-addressable.  Keys wider than 384 are
                                      accumulated in K segments of 384 (the tensor cores' truncating accumulator drifts ~1.2e-8 d |k|^2 in one
                                      accumulation: past the 1e-5 bar from d = 768 on); keys wider than 6144 go to FFMA */
int  b200p_lost_workspace_bytes(int n_images, int64_t total_patches, int64_t total_a, int d, int gram_impl, int64_t* out);
int  b200p_lost_batched(int device, const float* d_feats, int64_t row_stride, int d,
                        const b200p_lost_image_t* h_meta, int n_images, int k_patches,
                        float* d_A, int32_t* d_degree, int32_t* d_seed, float* d_box,
                        int32_t* d_status, void* d_workspace, int64_t workspace_bytes,
                        int gram_impl, void* stream);

/* Measurement aid: globaltimer (ns) trace of the last count-only b200p_lost_batched call: [Gram first CTA start, Gram last
 * CTA end, first finish CTA past its wait, last finish CTA end].  Synchronises the device. */
int  b200p_lost_last_trace(uint64_t* h_out4);
/* Measurement aid: globaltimer (ns) stamps of one finish CTA at its phase boundaries (lost.cu: g_fin_stamps) from the
 * last call; `image` selects the image the NEXT call traces.  Synchronises the device. */
int  b200p_lost_finish_trace(int image, uint64_t* h_out16);
/* Measurement aid: globaltimer (ns) stamps written by the last CTA of the most recent sample and sweep kernels on the
 * current device (select.cu: g_sel_stamps), then cleared.  Synchronises the device. */
int  b200p_select_last_trace(uint64_t* h_out16);

/* patch_scoring(M, threshold) of object_discovery.py:72-90 on a given n x n matrix (row stride lda):
 * d_degree[i] = #{j : (i != j ? max(A_ij, 0) : 0) > threshold}; d_sel = patches by ascending degree,
 * lowest index first among equal degrees (the reference's argsort(-degree, descending) made stable). */
int  b200p_lost_patch_scoring(int device, const float* d_A, int n, int64_t lda, float threshold,
                              int32_t* d_degree, int64_t* d_sel, void* stream);
/* detect_box of object_discovery.py:93-134 on a given correlation vector d_M [dim0*dim1]:
 * d_box = [xmin,ymin,xmax,ymax] in pixels (img_h / img_w <= 0: no clipping), d_feat_box =
 * [ymin,xmin,ymax,xmax] in feature cells (max exclusive), d_status = 1 if the seed is background. */
int  b200p_lost_detect_box(int device, const float* d_M, int dim0, int dim1, int seed, float scale0, float scale1,
                           int img_h, int img_w, float* d_box, int32_t* d_feat_box, int32_t* d_status, void* stream);

/* ---- host-buffer convenience entry points (the "e2e" path of bench.py) -------------- */
/* Complete SNIP mask build from HOST buffers: weights h_w [N], B gradient sets h_g[b] [N]
 * (flat, segment-concatenated, pinned or pageable), sparsity -> packed mask on the host.
 * H2D copies are double-buffered against the score kernels.  Synchronises. */
int  b200p_snip_mask_build_host(b200p_plan* plan, const float* h_w, const float* const* h_g,
                                int n_batches, uint64_t k, uint32_t* h_mask_out,
                                b200p_select_result_t* h_result);
int  b200p_magnitude_mask_build_host(b200p_plan* plan, const float* h_w, const uint32_t* h_old_mask,
                                     uint64_t k, uint32_t* h_mask_out,
                                     b200p_select_result_t* h_result);

#ifdef __cplusplus
}
#endif
#endif  /* B200PRUNE_H_ */
