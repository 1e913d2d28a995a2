/*
 * fa_b200.h -- C ABI of the B200-native (sm_100a) FlashAttention-forward engine.
 *
 * Drop-in boundary for the hot path of jeehun98/FlashAttention_Kernel_Project. The reference
 * has no library target: its operator surface is the parameter lists of its __global__ kernels
 * plus the main() conventions around them. Each entry point below names the reference interface
 * it replaces (paths relative to the reference repo).
 *
 * Conventions shared by every entry point
 *   - all tensor pointers are DEVICE pointers owned by the caller (the *_host variants take HOST
 *     pointers and perform the copies the reference's main() performs); nothing is retained;
 *   - row-major, contiguous; fp16 tensors are IEEE binary16 (`__half`), passed as `const void*`
 *     so this header needs no CUDA headers;
 *   - `stream` is a `cudaStream_t` (NULL = default stream); work is enqueued asynchronously;
 *   - return value: 0 on success, a negative FA_ERR_* for a rejected call, a positive
 *     `cudaError_t` for a CUDA runtime failure.  Nothing is printed and the process is never
 *     aborted (the reference's CHECK_CUDA macro exits; its kernels silently `return` on
 *     unsupported shapes, e.g. FlashAttention/flashattn_forward_wmma/flashattn_forward_wmma.cu:63);
 *   - there is NO CPU fallback: without an sm_100 device every compute call returns
 *     FA_ERR_NO_DEVICE.
 */
#ifndef FA_B200_H_
#define FA_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FA_OK 0
#define FA_ERR_BAD_SHAPE (-1)      /* non-positive extent, or product overflows the index range   */
#define FA_ERR_BAD_HEAD_DIM (-2)   /* D not in {64, 128} (fa_forward_cross: {16, 64, 128})        */
#define FA_ERR_MISALIGNED (-3)     /* pointer not 16-byte aligned / row pitch not a multiple of 16 B */
#define FA_ERR_NULL_POINTER (-4)
#define FA_ERR_NO_DEVICE (-5)      /* current device is not compute capability 10.x               */
#define FA_ERR_DRIVER_ENTRY (-6)   /* cuTensorMapEncodeTiled could not be resolved / failed       */
#define FA_ERR_BAD_ARGUMENT (-7)
#define FA_ERR_NCCL (-8)           /* fa_ring_forward: an NCCL call failed (or libnccl could not be resolved: -6) */

/* Library version: major*10000 + minor*100 + patch. */
int fa_version(void);
/*
 * Optional one-time setup for the CURRENT device: probes it, allocates the mapped debug page and the stream-ordered
 * memory pool of the attention kernel's tail workspace.  Every compute entry point does this lazily on first use; call
 * fa_init() once per device beforehand when the first compute call would otherwise happen inside a CUDA stream capture
 * (cudaHostAlloc and pool creation are not capturable).  Returns FA_OK or FA_ERR_NO_DEVICE.
 */
int fa_init(void);
/* Static string for a return code of this library (negative) or of CUDA (positive). */
const char* fa_strerror(int code);

/*
 * Non-causal attention forward, O = softmax(scale * Q K^T) V, per (b, h).
 *   Q, K, V : fp16 [B, H, N, D];  O : fp16 [B, H, N, D];  D in {64, 128};  any N >= 1.
 * Replaces: flashattn_forward_wmma_v5_cp_async_kernel(Q, K, V, O, BH, N, D, scale)
 *   (FlashAttention/flashattn_forward_memory_bound/flashattn_forward_wmma_v5_cp_async.cu:99-108)
 * and flashattn_forward_wmma_v4_kernel (FlashAttention/flashattn_forward_wmma/flashattn_forward_wmma_v4.cu:52-61);
 * the same math as the named comparators flashattn_2warp_nosmemW_online_softmax
 *   (flashattn_warp_spc/flashattn_streaming_16x32_2warp_softmax_v12i.cu:60-70).
 * B*H is the reference's flat `BH` batch; `scale` is explicit as in the reference (1/sqrt(D) there).
 */
int fa_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int N, int D,
               float scale, void* stream);

/* Same, fp32 output as the reference kernels write it (`float* O`, flashattn_forward_wmma.cu:49-58). */
int fa_forward_f32out(const void* Q, const void* K, const void* V, float* O, int B, int H, int N, int D,
                      float scale, void* stream);

/*
 * Attention with a query count different from the key count (decode / cross shapes), fp32 output:
 *   Q fp16 [B,H,Nq,D], K, V fp16 [B,H,Nkv,D], O fp32 [B,H,Nq,D];  D in {16, 64, 128}.
 * D = 16 covers the comparators BASELINE.json names at their native shapes:
 *   flashattn_2warp_nosmemW_online_softmax<16,16,4,32>(Q, K, V, O, seq_len, m_total, scale)
 *     (flashattn_warp_spc/flashattn_streaming_16x32_2warp_softmax_v12i.cu:60-70; B=1024, M=16, L=128)
 *   flashattn_streaming_16x16_kernel_mw_v7_5(Q, K, V, O, S, inv_sqrt_d)
 *     (flashattn_warp_spc_v2/flashattn_streaming_16x16_mw_v7_5_spinless_cpasync2stage.cu:84-91; one query per b: Nq = 1)
 * and the D = 16 streaming family (flashattn_warp_spc/flashattn_streaming_16x16_mw_v10.cu:104-112, with K as [B,L,16]).
 */
int fa_forward_cross(const void* Q, const void* K, const void* V, float* O, int B, int H, int Nq, int Nkv, int D,
                     float scale, void* stream);

/*
 * One step of a context-parallel (ring) attention: attends Q [B,H,Nq,D] to ONE K/V shard
 * [B,H,Nkv,D] and folds the result into the carried state
 *   O_acc fp32 [B,H,Nq,D] (un-normalised), m fp32 [B,H,Nq] (running row max of scale*QK^T),
 *   l fp32 [B,H,Nq] (running row sum),
 * with the merge rule the reference applies tile to tile
 *   (Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu:247-261).
 * is_first != 0: the incoming state is ignored (treated as empty).
 * is_last  != 0: the merged result is normalised and written to O_final (fp16 [B,H,Nq,D]) INSTEAD of
 *                the state -- (O_acc, m, l) are consumed and left as they were (no 4x-wider fp32 write
 *                on the step that ends the ring).  O_final may be NULL when is_last == 0.
 * The reference has no multi-GPU path; this entry exists for BASELINE config 4.
 */
int fa_forward_partial(const void* Q, const void* K, const void* V, float* O_acc, float* m, float* l,
                       int B, int H, int Nq, int Nkv, int D, float scale, int is_first, int is_last,
                       void* O_final, void* stream);

/*
 * Context-parallel (ring) attention over `world` ranks, one process per GPU (BASELINE config 4: B=1, H=32,
 * N=65536, D=128 over 2/4/8 GPUs).  Rank r holds rows [r*n, (r+1)*n) of Q, K, V (fp16 [B,H,n,D]) and receives the
 * same rows of O = softmax(scale * Q K_all^T) V_all.  The whole loop is enqueued from C++: `world` calls of the
 * fa_forward_partial kernel on `stream`, the K/V exchange as grouped ncclSend / ncclRecv on an internal
 * high-priority stream, tied together with events (no host synchronisation; CUDA-graph capturable).  K/V are
 * read-only, so every rank sends its OWN shard straight to the rank that needs it next (exchange x: to rank + x,
 * from rank - x) instead of passing shards along: the exchanges are independent, all of them are in flight from
 * the start (as many as the workspace has receive buffers), and step 0 -- the local shard -- needs none.
 *   nccl_comm : an ncclComm_t spanning the `world` ranks, this process being `rank` (made with
 *               ncclCommInitRank, or borrowed from the host framework's process group); the library resolves ncclSend/ncclRecv/ncclGroupStart/End from the
 *               libnccl.so.2 loaded in the process and links none itself.  May be NULL when world == 1.
 *   workspace : device memory owned by the caller, 256-byte aligned, at least fa_ring_workspace_bytes(..., depth)
 *               bytes: the carried fp32 (O_acc, m, l) + `depth` receive buffers of one K and one V shard.
 *               depth <= 0 or >= world-1 means world-1 buffers (everything prefetched); fewer buffers are reused
 *               in order, each exchange then waits for the step that last read its buffer.
 * The merge of the partial results is the reference's tile-to-tile (m, l, y) rule
 * (Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu:247-261); the reference has no multi-GPU path.
 * Every rank of the communicator must make the same call (same shapes) in the same order.
 */
size_t fa_ring_workspace_bytes(int B, int H, int n, int D, int world, int depth);
int fa_ring_forward(const void* Q, const void* K, const void* V, void* O, int B, int H, int n, int D, float scale,
                    int rank, int world, void* nccl_comm, void* workspace, size_t workspace_bytes, void* stream);

/*
 * The same ring with the exchange FUSED INTO THE ATTENTION KERNEL: peer_k[r] / peer_v[r] are addresses, valid in
 * THIS process, of rank r's K / V shard (peer mappings made with fa_ipc_open, or plain pointers when one process
 * drives several GPUs; entry `rank` is ignored).  The launch of step x attends to the shard of rank - x and, with the
 * two warps of every CTA that are otherwise idle, loads the shard of rank - (x + 1) straight out of its owner's
 * memory over NVLink / NVSwitch into the other of two local buffers (ld.global.cv on the peer mapping, 16 bytes per
 * lane, every CTA its 1 / grid slice).  One launch per step on the caller's stream and nothing else: no NCCL, no
 * second stream, no events, no SM taken from the attention kernel (NCCL's send/recv CTAs cannot share an SM with a
 * 227 KB attention CTA, and at 8 ranks its point-to-point exchange ran at ~70 GB/s per peer: 14.3 ms for config 4
 * against 7.7 ms of compute), and no rank waits for another inside the call.  (The copy engines would also do:
 * cudaMemcpyAsync from a mapping opened in the consumer's context measured 725 GB/s, 691 GB/s next to a running
 * attention kernel; the fused form needs no second stream and no events.)
 * Arithmetic, visiting order and workspace layout are those of fa_ring_forward (two receive buffers are enough):
 * the two return bit-identical results.
 * The caller orders the ranks around the call: every rank's K / V must be complete before any rank starts, and
 * unmodified until every rank's call has finished (one barrier on each side).  fa_enable_peer_access(d) enables
 * direct access from the current device to device d (once per pair; without it the driver stages peer copies
 * through host memory).
 */
int fa_enable_peer_access(int peer_device);
/*
 * Peer mappings between the one-process-per-GPU ranks (thin wrappers of CUDA IPC; no framework needed):
 *   fa_ipc_export(ptr, handle, &offset)  in the OWNER process: the 64-byte handle of the allocation that contains
 *                                        `ptr` and ptr's byte offset inside it (send both to the other ranks);
 *   fa_ipc_open(handle, &base)           in a CONSUMER process, with the consuming GPU current: maps the allocation
 *                                        with peer access over NVLink enabled; the shard is at base + offset.  Open a
 *                                        given handle once per process and share the mapping (K and V of one rank
 *                                        often live in one allocation);
 *   fa_ipc_close(base)                   unmaps; the owner may free the allocation afterwards.
 * The mapping must be opened in the CONSUMER's context: a mapping opened on the owner's device index (what a
 * framework's tensor sharing does) can be copied from by the driver (staged: 25 GB/s measured, against 725 GB/s for
 * a consumer-context mapping) but faults when a kernel of another GPU loads from it.
 */
int fa_ipc_export(const void* ptr, unsigned char* handle /* [64] */, unsigned long long* offset);
int fa_ipc_open(const unsigned char* handle /* [64] */, void** base);
int fa_ipc_close(void* base);
int fa_ring_forward_pull(const void* Q, const void* K, const void* V, void* O, int B, int H, int n, int D, float scale,
                         int rank, int world, const void* const* peer_k, const void* const* peer_v, void* workspace,
                         size_t workspace_bytes, void* stream);

/*
 * C = ReLU(A*B + bias[col]) + residual   (epilogue order: bias -> ReLU -> residual).
 *   A fp16 [M,K] row-major; Bm fp16 [K,N] row-major (NOT transposed); bias fp32 [N] or NULL;
 *   residual fp32 [M,N] or NULL; C fp32 [M,N].  K % 8 == 0 and N % 8 == 0 (16-byte row pitch).
 * Replaces: wmma_gemm_fused_kernel(A, B, bias, residual, C, M, N, K)
 *   (GEMM/tensorcore_gemm_fused_epilogue/tensorcore_gemm_fused_epilogue.cu:53-58, epilogue :109-126).
 */
int gemm_fused_epilogue(const void* A, const void* Bm, const float* bias, const float* residual, float* C,
                        int M, int N, int K, void* stream);

/*
 * The same kernel with a selectable activation between bias and residual: C = act(A*B + bias[col]) + residual.
 * FA_ACT_RELU is gemm_fused_epilogue; identity / exact (erf) GELU / SiLU complete the epilogue surface -- the
 * reference implements ReLU only (tensorcore_gemm_fused_epilogue.cu:117; gemm_fused_epilogue.cu:35-37).
 */
#define FA_ACT_IDENTITY 0
#define FA_ACT_RELU 1
#define FA_ACT_GELU 2 /* 0.5 x (1 + erf(x / sqrt 2)) */
#define FA_ACT_SILU 3 /* x / (1 + exp(-x)) */
int gemm_fused_epilogue_act(const void* A, const void* Bm, const float* bias, const float* residual, float* C,
                            int M, int N, int K, int act, void* stream);

/*
 * The fp32-operand sibling: C = act(A*B + bias[col]) + residual with A fp32 [M,K], Bm fp32 [K,N] (row-major, not
 * transposed), fp32 bias / residual / C; K % 4 == 0 and N % 4 == 0.
 * Replaces: gemm_tiled_fused_kernel(A, B, bias, residual, C, M, N, K)
 *   (GEMM/gemm_fused_epilogue/gemm_fused_epilogue.cu:44-111: 32 x 32 shared-memory tiles, scalar fp32 FMA, the same
 *    bias -> ReLU -> residual epilogue, :88-110), act = FA_ACT_RELU.
 * Runs on the tensor cores at fp32-level accuracy: every operand is split into two tf32 numbers, x = hi + lo with
 * hi = x rounded to tf32 and lo = x - hi (exact), and the product is computed as A_lo*B_hi + A_hi*B_lo + A_hi*B_hi
 * ("3xTF32"; the dropped lo*lo term is below 2^-21 relative) by the same persistent CTA-pair kernel as
 * gemm_fused_epilogue (tcgen05.mma kind::tf32, cta_group::2, TMA epilogue), the large term and the two cross terms
 * in separate fp32 accumulators in TMEM (the tensor core's accumulation truncates: fewer steps on the large sum),
 * added in the epilogue.  Measured max-abs error 1.4e-4 at the reference's 1024^3 uniform(-1,1) fixture, where a
 * plain tf32 GEMM is 1.3e-2 off (limit 2e-3).  A short pre-pass writes the four split operands (B transposed to [N,K] on the way) into scratch taken
 * from the library's stream-ordered pool for the duration of the call: 2 x (M*K + N*K) floats.
 */
int gemm_fused_epilogue_f32(const float* A, const float* Bm, const float* bias, const float* residual, float* C,
                            int M, int N, int K, int act, void* stream);

/*
 * Batched plain GEMM with a scale: C[z] = alpha * A[z] * op(B[z]) for z in [0, batch), no bias / activation.
 *   A fp16 [batch,M,K]; Bm fp16 [batch,K,N] if trans_b == 0, fp16 [batch,N,K] if trans_b != 0 (C = alpha*A*B^T);
 *   C fp32 [batch,M,N]; all row-major and densely packed; K % 8 == 0 and N % 8 == 0.
 * Replaces: wmma_batched_gemm_kernel(A, B, C, batch, M, N, K)            (trans_b = 0, alpha = 1)
 *             (GEMM/gemm_tensorcore_batched/gemm_4_9_tensorcore_batched.cu:42-46)
 *           QK_wmma_kernel(Q, K, Scores, BH, N, D, scale)                 (trans_b = 1: A = Q, Bm = K, M = N = seq, K = D, alpha = scale)
 *             (GEMM/QK Tensor Core/qk_wmma_5_1_2.cu:32-38)
 *           wmma_gemm_tensorcore_kernel(A, B, C, M, N, K)                 (batch = 1; the QK^T and PV stages of
 *             FlashAttention/flashattn_forward_memory_bound/flashattn_stage_latency_breakdown.cu:58-65)
 */
int gemm_batched(const void* A, const void* Bm, float* C, int batch, int M, int N, int K, int trans_b, float alpha,
                 void* stream);

/*
 * gemm_batched on strided operands (BLAS-style leading dimensions and batch strides, all in ELEMENTS; 0 = densely
 * packed): row pitches lda (A rows of K), ldb (B rows: N if trans_b == 0, K otherwise), ldc (C rows of N) and the
 * distances stride_a / stride_b / stride_c between consecutive batch entries.  Pitches and strides must keep rows
 * 16-byte aligned (multiples of 8 fp16 / 4 fp32 elements).  The reference hard-codes the packed case
 * (strideA = M*K, strideB = K*N, strideC = M*N, leading dimensions K, N, N:
 * GEMM/gemm_tensorcore_batched/gemm_4_9_tensorcore_batched.cu:58-60,77-78,85); with strides the same kernel runs
 * on sub-matrices and interleaved batches, e.g. one head of a [B,N,H,D] tensor, without a copy.
 */
int gemm_batched_strided(const void* A, const void* Bm, float* C, int batch, int M, int N, int K, long long lda,
                         long long ldb, long long ldc, long long stride_a, long long stride_b, long long stride_c,
                         int trans_b, float alpha, void* stream);

/*
 * Row softmax, P[r, :] = softmax(X[r, :]) with max subtraction; X, P fp32 [rows, N]; in place allowed.
 * Replaces: softmax_naive_baseline_kernel(scores, probs, num_rows, N)
 *   (Softmax micro kernel/softmax_naive_baseline/softmax_naive_baseline.cu:56-61).
 */
int softmax_rows(const float* X, float* P, int rows, int N, void* stream);

/* Same, fp16 output (the A operand of a following P*V GEMM); X and P must not overlap. */
int softmax_rows_f16out(const float* X, void* P, int rows, int N, void* stream);

/*
 * Attention through MATERIALISED scores, three launches on `stream`:
 *   S = scale * Q K^T (gemm_batched, trans_b) -> P = softmax_rows_f16out(S) -> O = P V (gemm_batched).
 *   Q, K, V fp16 [B,H,N,D]; O fp32 [B,H,N,D]; workspaces owned by the caller: S fp32 [B*H,N,N], P fp16 [B*H,N,N].
 *   N % 8 == 0 and D % 8 == 0 (16-byte row pitches); any D.
 * This is the reference's unfused pipeline -- QK_wmma_kernel (GEMM/QK Tensor Core/qk_wmma_5_1_2.cu:32-38), a row
 * softmax, a second WMMA GEMM -- whose per-stage cost its latency study measures
 * (FlashAttention/flashattn_forward_memory_bound/flashattn_stage_latency_breakdown.cu:212-424).  It moves
 * 6*N*N bytes per (b,h) through HBM that fa_forward never writes; it exists as the independent second
 * implementation the fused kernel is checked against at full size, and to reproduce that study on a B200.
 */
int fa_forward_unfused(const void* Q, const void* K, const void* V, float* O, float* S, void* P, int B, int H, int N,
                       int D, float scale, void* stream);

/*
 * Host-buffer variants: the flow of the reference's main() (cudaMalloc -> H2D -> launch -> D2H,
 * e.g. flashattn_streaming_16x32_2warp_softmax_v12i.cu:325-368) behind one call.  All pointers
 * are HOST pointers (pinned memory makes the copies asynchronous); device scratch is cached
 * inside the library per device and grown on demand.  The call returns after the result is in
 * the host buffer.  The batch is cut into up to 16 chunks of independent units (copies of at least 2 MB) whose H2D
 * copy, kernel and D2H copy overlap on three internal streams (environment FA_HOST_CHUNKS = 1..64 overrides the
 * bound, FA_HOST_CHUNK_MB the smallest copy; config 3 measures 16.2-17.1 ms for 8..64 chunks on one GPU, i.e. the call
 * is bound by the 805 MB host-to-device copy).  Calls on one device are serialised (they share the scratch);
 * calls on different devices run concurrently.  fa_host_release() frees the cached scratch of the current device.
 */
int fa_forward_host(const void* Q, const void* K, const void* V, void* O, int B, int H, int N, int D,
                    float scale);
int gemm_fused_epilogue_host(const void* A, const void* Bm, const float* bias, const float* residual,
                             float* C, int M, int N, int K);
int softmax_rows_host(const float* X, float* P, int rows, int N);
int fa_host_release(void);

/*
 * Hang diagnostics: every mbarrier wait in the kernels is bounded (30 s of wall clock by default, far above any
 * legitimate wait even when the context is time-sliced or replayed under a profiler; -DFA_WAIT_TIMEOUT_NS=<ns> at
 * build time changes the bound, 0 removes it); on expiry the kernel
 * records where it was stuck and traps.  fa_last_debug copies up to `n` 32-bit words of that
 * record (word0 = 0xDEAD0001 if a record exists; site, blockIdx.x, blockIdx.y, threadIdx.x,
 * parity, tag follow) and returns the number of words written.  The B200-native counterpart of
 * the reference's "[STUCK]" printf hang hunt
 * (flashattn_warp_spc_v2/flashattn_streaming_16x16_mw_v7_2_dbg_hanghunt.cu:163,231-233).
 */
int fa_last_debug(unsigned int* words, int n);

/* Number of kernels this library has launched in this process (all entry points). */
unsigned long long fa_launch_count(void);

/*
 * Work decomposition of the attention kernel on a device with `sms` SMs (host arithmetic only, no device work).
 * The reference's grid is one CTA per (query block, b*h) (FlashAttention/flashattn_forward_wmma/flashattn_forward_wmma_v4.cu:452);
 * here an ITEM is one 256-row query block of one (b,h), CTAs of the full waves own one item each, and the items of
 * the partial last wave are cut along K/V into equal spans of 128-key tiles, one CTA per span ("stream-K" tail; the
 * parts of an item are merged with the (m, l, O) rule of Streaming_FlashAttention_Forward_Kernel/flashattn_streaming_16x16.cu:247-261).
 * A last wave that fills at most half the SMs can instead be cut along QUERY ROWS: two CTAs per item, one 128-row
 * tile each, every row computed exactly as in the whole item and nothing to merge (the reference's own grid has one
 * CTA per 64-row block, flashattn_forward_wmma_v4.cu:452).
 * Writes up to max_segments rows {cta, item, kv_tile_begin, kv_tile_end, slot, tail_item} into `out` (may be NULL)
 * and returns the number of segments, or a negative FA_ERR_*.  slot: -1 = whole item; 0 / 1 = a K/V part (which of
 * the CTA's two workspace slots holds it); -2 / -3 = rows [0,128) / [128,256) of the item against all its key tiles.
 * Environment FA_TAIL_SPLIT (read once): 0 never cuts, 1 (default) cuts where measurements say it pays (K/V cut for
 * grids that fill less than half the SMs, row cut for small last waves), 2 cuts the partial last wave of any grid
 * along K/V, 3 applies only the row cut.  fa_set_tail_split(0 ... 3) overrides the
 * environment for the process, fa_set_tail_split(-1) hands the decision back to it; the previous setting is returned.  A cut item
 * is summed in a different order than a whole one (its parts are merged with the (m, l, O) rule), so a (b,h) slice
 * is bit-identical to the same slice of a larger batch only when both are decomposed alike -- e.g. with the cut off.
 */
int fa_plan_segments(int B, int H, int Nq, int Nkv, int sms, int* out, int max_segments);
int fa_set_tail_split(int mode);

/* Tensor-map cache counters (the cache is bounded and mutex-protected); returns its capacity. */
int fa_map_cache_stats(unsigned long long* hits, unsigned long long* misses);

#ifdef __cplusplus
}
#endif
#endif /* FA_B200_H_ */
