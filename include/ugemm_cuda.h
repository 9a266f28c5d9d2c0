/* ugemm_cuda.h -- Blackwell (sm_100a) SGEMM backend for ugemm: the drop-in C ABI.
 *
 * This is the ONLY header a ugemm user includes to use the CUDA backend.  It sits next to the
 * reference's sgemm_avx256.h / sgemm_sse.h / sgemm_ocl*.h / sgemm_gl*.h backends and keeps their
 * conventions: BLAS-style 14-argument entry points returning void, caller-owned host buffers,
 * blocking calls, process-global state bracketed by init/finish, errors reported out of band.
 * Every symbol is plain C (extern "C", pointers and sizes only); the implementation is
 * libugemm_cuda.so (ugemm_b200/csrc/ *.cu, hand-written sm_100a kernels, no CPU fallback).
 *
 * Semantics (identical to the reference, ugemm.h:305-409):
 *   C <- alpha * op(A) * op(B) + beta * C        fp32 storage, fp32-accurate arithmetic
 *   major  'R' row-major | 'C' column-major      trans 'N' | 'T' (lower case accepted)
 *   row-major:  op(A)(m,k) = transA=='N' ? A[k + m*lda] : A[m + k*lda]
 *               op(B)(k,n) = transB=='N' ? B[n + k*ldb] : B[k + n*ldb]     C(m,n) = C[n + m*ldc]
 *   column-major is the mirror image (ugemm.h:359-409).
 *   beta == 0   C is overwritten and never read (as sgemm_avx / sgemm_c / sgemm_sse do:
 *               sgemm_avx256.h:324-330, gemm_cpu.h:119-124); NaN/Inf in C do not propagate.
 *   alpha == 0 or K == 0   C <- beta*C over the M x N region in the GIVEN major (the reference's
 *               fast paths index column-major regardless -- sgemm_avx256.h:415-430 -- a bug this
 *               backend does not reproduce).
 *   Elements of C outside the M x N region (ld padding) are never written.
 *   Dimensions and leading dimensions stay `int` for drop-in compatibility; all internal
 *   offsets are 64-bit (32768 x 32768 operands are supported).
 *   Non-finite inputs: K2 (plain fp32) propagates NaN / +-Inf exactly like the reference's loops.  K1 (3xTF32) marks the same
 *   entries of C non-finite, but an entry the reference reports as +-Inf may come out as NaN: the error-compensation term
 *   a_big * b_small is Inf * 0 whenever the other operand is exactly representable in TF32 (tests/test_parity_gpu.py::
 *   test_non_finite_inputs).  Entries whose inputs are all finite are unaffected.
 *
 * Threading: the backend is process-global and lives on ONE device (like the reference's single cl context, ocl.h:141-193).
 * The host-pointer entry points serialise on an internal lock (they share the staging arena).  The *_dev entry points may be
 * called from several host threads on their own streams; each call binds the calling thread to the backend's device.
 * sgemm_cuda_last_kernel / _last_repacked / _launch_count / _last_error report the most recent call of ANY thread.
 */
#ifndef UGEMM_CUDA_H
#define UGEMM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Kernel selection for the *_dev entry point and for reports. */
enum {
	UGEMM_MODE_AUTO   = 0, /* rule-based: K1 iff eligible (see sgemm_cuda_k1_eligible), else K2 */
	UGEMM_MODE_3XTF32 = 1, /* K1: 3xTF32 error-compensated tcgen05/TMEM/TMA kernel; ineligible => error */
	UGEMM_MODE_SIMT   = 2  /* K2: register-blocked FFMA kernel, plain fp32 */
};

/* ---- lifecycle: replaces sgemm_ocl_init / sgemm_ocl_finish (sgemm_ocl2.h:158-165, :220-224) and the
 * device setup in ocl.h:141-249 (oclSetup / oclKernel / oclKernelArgs) and gpgpu_gl4.h:142-169 (coInit).
 * `device` = CUDA ordinal (-1: $UGEMM_CUDA_DEVICE or 0).  `arena_bytes` = initial size of the device
 * staging arena used by the host-pointer entry points (0: grow on demand), like the single cl buffer the
 * OpenCL backend sizes up front.  Returns 0 on success; on failure returns non-zero and sets last_error.
 * Calling it again is a no-op for the same device (the arena grows if asked) and an ERROR for another device: call
 * sgemm_cuda_finish first.  $UGEMM_K1_FLAGS (tuning / A-B switches of K1) is read here; its ablation bits, which make
 * results wrong, are rejected unless $UGEMM_K1_ABLATION=1. */
int  sgemm_cuda_init(int device, size_t arena_bytes);
void sgemm_cuda_finish(void);

/* ---- drop-in SGEMM, HOST pointers, blocking: H2D -> kernel -> D2H exactly like sgemm_ocl
 * (sgemm_ocl2.h:166-218).  Same signature as sgemm_cpu (ugemm.h:287-301), sgemm_c (gemm_cpu.h:284-298),
 * sgemm_avx (sgemm_avx256.h:392-405), sgemm_sse (sgemm_sse.h:365-379), i.e. usable as the `uut`
 * function pointer of test_sgemm (check_sgemm.c:96-103).  Lazily calls sgemm_cuda_init(-1, 0). */
void sgemm_cuda(char major, char transA, char transB, int M, int N, int K, float alpha,
                const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc);
/* forced-kernel variants so a harness can list both kernels as separate `uut` rows */
void sgemm_cuda_3xtf32(char major, char transA, char transB, int M, int N, int K, float alpha,
                       const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc);
void sgemm_cuda_simt(char major, char transA, char transB, int M, int N, int K, float alpha,
                     const float *A, int lda, const float *B, int ldb, float beta, float *C, int ldc);

/* Reproducibility: a call is deterministic -- the same arguments (shape, leading dimensions, pointer alignment) on the same device
 * with the same SM limit give the same bits, stream-K tail and serpentine K included (no atomics on data).  Two DIFFERENT ways of
 * computing the same product (the pipelined host-pointer path, which multiplies row panels; the sharded drivers; another SM limit)
 * order the K sum differently and agree to round-off (~1e-6 relative), not bit for bit. */

/* ---- the timed path: DEVICE pointers, asynchronous on `stream` (a cudaStream_t; NULL = the backend's
 * own non-blocking stream -- pass cudaStreamLegacy (0x1) to mean CUDA's legacy default stream).  Replaces the body of sgemm_ocl after its uploads (kernel launch, sgemm_ocl2.h:201-215),
 * including the work of the separate `transpose` kernel (sgemm_ocl2.h:95-128,180-199): transposes are
 * folded into the operand loads.  Returns 0 on success. */
int sgemm_cuda_dev(int mode, void *stream, char major, char transA, char transB, int M, int N, int K,
                   float alpha, const float *dA, int lda, const float *dB, int ldb,
                   float beta, float *dC, int ldc);

/* ---- strided batch: `batch` problems of one shape in ONE launch; instance b uses A + b*strideA, B + b*strideB,
 * C + b*strideC (strides in elements).  This is the stacked-instance layout of the reference's test_sgemm (11 instances,
 * a is (11*M) x lda etc., check_sgemm.c:111-124,242-246), which the reference walks one call at a time.  K1 needs
 * strideA and strideB to be multiples of 4 in addition to the usual rule; otherwise K2.  No operand repacking here.
 * Strides must be non-negative and strideC must keep the instances of C apart; with alpha == 0 or K == 0 A and B are not read. */
void sgemm_cuda_batched(char major, char transA, char transB, int M, int N, int K, float alpha,
                        const float *A, int lda, long long strideA, const float *B, int ldb, long long strideB,
                        float beta, float *C, int ldc, long long strideC, int batch);             /* host pointers, blocking */
int  sgemm_cuda_batched_dev(int mode, void *stream, char major, char transA, char transB, int M, int N, int K, float alpha,
                            const float *dA, int lda, long long strideA, const float *dB, int ldb, long long strideB,
                            float beta, float *dC, int ldc, long long strideC, int batch);        /* device pointers, async */

/* 1 if UGEMM_MODE_AUTO would pick K1 DIRECTLY for this problem: A, B 16-byte aligned, lda and ldb multiples
 * of 4 (TMA global-stride rule), K >= 32, and M,N >= 128 (at least one full tile of tensor work) -- or one of M, N >= 128,
 * the other >= 48 (>= 8 when K >= 512) and M*N*K >= 2^26 (a skinny but large product: K1's padded tile still beats the FFMA kernel).
 * The complete auto rule: (1) that -> K1; (2) else, if M,N >= 256 and K >= 64, the operand(s) TMA cannot take are
 * first copied to a stream-ordered scratch buffer with an aligned leading dimension (one HBM pass, reported by
 * sgemm_cuda_last_repacked) and K1 runs on the copy; (3) else K2.  Forced modes never repack. */
int sgemm_cuda_k1_eligible(char major, char transA, char transB, int M, int N, int K,
                           const float *dA, int lda, const float *dB, int ldb, const float *dC, int ldc);

/* Time `iters` back-to-back launches (after `warmup` untimed ones) with CUDA events on the backend's stream:
 * one event pair per launch, all enqueued without host synchronisation in between.  *ms_avg / *ms_min = mean /
 * best single-launch duration, *ms_total = first start to last end (any may be NULL).  Returns 0 on success. */
int sgemm_cuda_time_dev(int mode, int iters, int warmup, char major, char transA, char transB,
                        int M, int N, int K, float alpha, const float *dA, int lda,
                        const float *dB, int ldb, float beta, float *dC, int ldc,
                        float *ms_avg, float *ms_min, float *ms_total);

/* ---- errors: entry points stay void like the reference's; failures are sticky and queryable
 * (the reference only printf()s, ocl.h:92,235-239).  NULL when no error is pending. */
const char *sgemm_cuda_last_error(void);
void        sgemm_cuda_clear_error(void);

/* ---- introspection used by the harnesses / bench */
int                sgemm_cuda_last_repacked(void);  /* 1 if the last auto launch went through rule (2) above */
int                sgemm_cuda_last_kernel(void);   /* UGEMM_MODE_3XTF32 or UGEMM_MODE_SIMT of the last GEMM launch, 0 if none */
unsigned long long sgemm_cuda_launch_count(void);  /* number of GEMM/fill/scale kernels launched by this library so far */
int  ugemm_cuda_device_info(int *sm_count, int *sm_clock_khz, size_t *hbm_bytes, char *name, int name_len);
/* Tunables of K1 (negative = keep).  kc_blocks: number of 32-wide k-blocks accumulated inside TMEM
 * before the partial sums are promoted to fp32 registers with round-to-nearest adds;
 * split: 0 = truncation split, raw tile is the "big" operand; 1 = round-to-nearest split, big rewritten.
 * cta_group: 1 or 2 CTAs per MMA, 0 = by problem size (pairs once >= ~3/4 of the SM pairs have a 256x256 tile). */
void sgemm_cuda_set_k1_tuning(int kc_blocks, int split, int cta_group);
/* Which implementation of K1 runs: 0 = TS (op(A) in tensor memory, 64-column accumulator slices; the default),
 * 1 = SS (round 1: both operands from shared memory, 2 x 256-column accumulators).  Same results within the gate;
 * kept selectable for A/B measurements and so that the tests exercise both.  The
 * round-to-nearest split experiment (split = 1) always runs on SS. */
void sgemm_cuda_set_k1_variant(int variant);

/* The schedule K1 would use for a dense M x N x K problem (`batch` instances) on `sm_count` SMs (0 = 148) with the current tuning,
 * as pure host arithmetic (no GPU): plan12[0..11] = cta_group, tile_m, tile_n, tiles_m, tiles_n, k-blocks per tile, promotion
 * interval in k-blocks, whole tiles, stream-K tail tiles, promotion chunks per tile, chunks per tail range, work items.  Work items
 * below plan12[7] are whole tiles; the others are chunk ranges of the tail, each of up to two segments:
 * sgemm_cuda_k1_plan_item gives segment h (0 / 1) of an item as out4 = tile, first k-block, end k-block, workspace slot (-1: a whole
 * tile; end <= first: no such segment) -- the very function the kernel's roles decode their work with.  For tests and tools. */
int  sgemm_cuda_k1_plan(int M, int N, int K, int batch, int sm_count, int *plan12);
int  sgemm_cuda_k1_plan_item(const int *plan12, int item, int h, int *out4);

/* Cap the number of SMs K1's persistent grid occupies (0 = all).  Used by the sharded driver while NCCL panel
 * broadcasts are in flight: a persistent CTA fills an SM's registers and shared memory, so a few SMs are left
 * free for the collective's own CTAs instead of serialising the transfer behind the GEMM. */
void sgemm_cuda_set_sm_limit(int sms);

/* ---- memory helpers (replace oclKernelArgs/oclWrite/oclRead buffer plumbing, ocl.h:227-293) */
void *ugemm_cuda_malloc(size_t bytes);            /* device memory */
void  ugemm_cuda_free(void *dptr);
void *ugemm_cuda_malloc_host(size_t bytes);       /* pinned host memory (fast H2D/D2H for sgemm_cuda) */
void  ugemm_cuda_free_host(void *hptr);
int   ugemm_cuda_memcpy_h2d(void *dst, const void *src, size_t bytes);
int   ugemm_cuda_memcpy_d2h(void *dst, const void *src, size_t bytes);
int   ugemm_cuda_sync(void);
/* stream-ordered copy between any two pointers of the unified address space (device<->device across GPUs included) */
int   ugemm_cuda_memcpy_async(void *dst, const void *src, size_t bytes, void *stream);
/* Peer-to-peer plumbing for the sharded driver (one process per GPU): export a ugemm_cuda_malloc'ed buffer as a
 * 64-byte CUDA IPC handle, map a peer's handle into this process (peer access enabled lazily), unmap it.  Copies
 * from a mapped peer buffer run on the copy engines over NVLink and take no SMs from the GEMM. */
int   ugemm_cuda_ipc_export(void *dptr, void *handle64);
void *ugemm_cuda_ipc_import(const void *handle64);
int   ugemm_cuda_ipc_close(void *mapped);

/* ---- single-process multi-GPU SGEMM (one host thread drives every GPU of the box; the C host program's way to the
 * sharded path -- bench.py's one-process-per-GPU twin is ugemm_b200/dist.py).  The reference has no multi-device path:
 * its OpenCL backend owns one device and one queue (ocl.h:141-193); this extends the uut signature (check_sgemm.c:96-103)
 * by a process grid.  C is cut into pr x pc blocks, GPU i*pc + j owns block (i, j) and receives A row-panel i and B
 * column-panel j; K is not split, so there is one exchange step and no reduction.  Panels travel slab by slab along K as a
 * pipelined relay on the copy engines over NVLink (caller's buffer -> GPU (i,0) -> (i,1) ...; B down the columns), the
 * product of slab t (beta = 1 after the first) runs while slab t+1 is in flight (overlap != 0), or after the whole
 * distribution (overlap == 0, which separates distribution time from compute time; overlap == 2 also cuts K into slabs on
 * a 1 x 1 grid, which is only useful for testing the slab arithmetic on one GPU).
 *   sgemm_cuda_mgpu_init(n): GPUs 0..n-1, peer access all-to-all, per-GPU streams and a cached arena; 0 = OK.
 *   sgemm_cuda_mgpu(...):    A, B, C are host pointers (pinned for speed) or device pointers of ANY GPU (unified
 *                            addressing); same argument checks, quick returns and alpha/beta/ld semantics as sgemm_cuda;
 *                            blocking; C's ld padding is never written; pr * pc <= n.  Returns 0 / 1 (sticky error).
 *                            A block whose source already lives on the GPU that needs it is used in place (no copy): with
 *                            the operands on GPU 0, GPU 0 multiplies at once and only serves its peers.
 *   timings_ms (optional, 5 floats): [0] host wall clock of the call, [1] max over GPUs of start -> C block written back,
 *                            [2] max over GPUs of start -> last panel slab landed, [3] max over GPUs of first product
 *                            start -> last product end, [4] max over GPUs of start -> last product end (the metric of
 *                            BASELINE config 5: distribution included, gather of C excluded). */
int  sgemm_cuda_mgpu_init(int ngpus);
void sgemm_cuda_mgpu_finish(void);
int  sgemm_cuda_mgpu_count(void);     /* GPUs initialised by sgemm_cuda_mgpu_init, 0 if none */
/* the partition sgemm_cuda_mgpu uses for a row-major M x N x K problem on a pr x pc grid: GPU (i, j) owns rows
 * [i*block_rows, ...) x columns [j*block_cols, ...) of C (the last blocks may be short or empty); K is cut into k_slabs slabs of
 * slab_width (the last may be short).  Pure host arithmetic, needs no GPU; 0 = OK, 1 = bad arguments. */
int  sgemm_cuda_mgpu_plan(int M, int N, int K, int pr, int pc, int overlap, int *block_rows, int *block_cols, int *k_slabs, int *slab_width);
int  ugemm_cuda_device_count(void);   /* CUDA devices visible to this process (0 without a driver); never an error */
int  sgemm_cuda_mgpu(char major, char transA, char transB, int M, int N, int K, float alpha, const float *A, int lda,
                     const float *B, int ldb, float beta, float *C, int ldc, int pr, int pc, int overlap, float *timings_ms);
/* the same call under the init / run / finish naming of the other backends' macro sets (sgemm_test.c:19-33) */
int  sgemm_cuda_mgpu_run(char major, char transA, char transB, int M, int N, int K, float alpha, const float *A, int lda,
                         const float *B, int ldb, float beta, float *C, int ldc, int pr, int pc, int overlap, float *timings_ms);

/* ---- sharded SGEMM, ONE PROCESS PER GPU (SURVEY section 8 e; csrc/shard.cu): the multi-process twin of sgemm_cuda_mgpu.
 * Row-major NN C = A * B on a pr x pc grid of C blocks (1x1, 2x1, 2x2, 2x4 for 1/2/4/8 ranks); rank r = i*pc + j owns block (i, j).
 * K is cut into slabs; slab t of A row-panel i starts on rank (i, t*pc/L), slab t of B column-panel j on rank (t*pr/L, j)
 * (owner-rooted placement).  Slabs travel over NVLink by NCCL broadcast inside grid-row / grid-column communicators
 * (transport 0: ncclCommInitRank + ncclCommSplit + ncclBroadcast; NCCL is loaded with dlopen) or by copy-engine peer pulls
 * from CUDA-IPC-mapped allocations (transport 1; falls back to 0 on every rank when one rank has no peer path).  The product of
 * slab t overlaps the transfer of the slabs behind it.  The reference owns one device (ocl.h:141-193); nothing there to replace.
 * The host program brings rendezvous only: rank 0 calls sgemm_cuda_shard_unique_id and hands the 128 bytes to every rank.
 * Every function returns 0 on success (message in sgemm_cuda_last_error); all ranks must make the same calls in the same order.
 *   _plan / _owners   pure host arithmetic (no GPU): the partition, who owns slab t and where it sits in the owner's allocation
 *   _init             communicators, streams, the owned-slab / received-slab / C-block allocations on the current device
 *   _generate         every rank synthesises the slabs it owns as windows of the global ugemm_fill_uniform streams
 *   _run              `warmup` + `steps` steps (distribute != 0: every slab travels in every step); *ms_total = this rank's
 *                     CUDA-event time of the `steps` timed steps, taken between two barriers
 *   _run_host         the same end to end and pipelined: owned slabs start in pinned host memory (H2D, broadcast and products
 *                     overlap slab by slab; the last slab's product runs in row panels whose C rows go down at once); host wall clock
 *   _allreduce        max (op 0) / sum (op 1) of one float over the ranks
 *   _block            device pointer and global window of this rank's C block */
int  sgemm_cuda_shard_plan(int world, int rank, int M, int N, int K, int *pr, int *pc, int *slabs, int *slab_width, int *block_rows, int *block_cols);
int  sgemm_cuda_shard_owners(int world, int rank, int M, int N, int K, int slab, int *a_owner_rank, int *b_owner_rank, long long *a_offset, long long *b_offset);
int  sgemm_cuda_shard_unique_id(unsigned char *id128);
int  sgemm_cuda_shard_init(int rank, int world, const unsigned char *id128, int M, int N, int K, int transport);
void sgemm_cuda_shard_finish(void);
int  sgemm_cuda_shard_transport(void);   /* transport in use (0 NCCL broadcast, 1 peer pull), -1 if not initialised */
int  sgemm_cuda_shard_generate(unsigned long long seed_a, unsigned long long seed_b, float lo, float hi);
int  sgemm_cuda_shard_run(int distribute, int steps, int warmup, float *ms_total);
int  sgemm_cuda_shard_allreduce(float *value, int op);
int  sgemm_cuda_shard_block(float **d_c, int *rows, int *cols, int *row0, int *col0);
int  sgemm_cuda_shard_host_buffers(float **h_own, long long *own_floats, float **h_c, long long *c_floats);
int  sgemm_cuda_shard_download_owned(void);
int  sgemm_cuda_shard_run_host(int steps, int warmup, float *ms_total, long long *h2d_bytes_per_step, long long *d2h_bytes_per_step);
/* the host-link floor of _run_host on this box: the same bytes up and down at once, no broadcast, no product; host wall clock */
int  sgemm_cuda_shard_copy_floor(int steps, float *ms_total);

/* ---- counter-based uniform stream, identical on host and device (so a 32768^2 operand can be generated
 * on the GPU and any sampled row regenerated on the host for verification):
 *   x[i] = fmaf(hi-lo, (splitmix64(seed*0x9E3779B97F4A7C15 + i) >> 40) * 2^-24, lo)
 * Plays the role of random_matrix (check_sgemm.c:47-54) with an explicit seed. */
void ugemm_fill_uniform_host(float *x, size_t n, uint64_t seed, float lo, float hi);
int  ugemm_fill_uniform_dev(float *dx, size_t n, uint64_t seed, float lo, float hi, void *stream);

/* 2-D window variants: dst[r*ld + c] = element (offset + r*gld + c) of stream `seed`, i.e. the rows x cols window
 * starting at flat index `offset` of a row-major matrix with leading dimension gld.  Lets every GPU of the sharded
 * driver generate exactly its own panel of one global synthetic matrix, and lets the host regenerate any slab of
 * it for verification. */
void ugemm_fill_uniform_host_2d(float *x, size_t rows, size_t cols, size_t ld, uint64_t seed, uint64_t offset,
                                uint64_t gld, float lo, float hi);
int  ugemm_fill_uniform_dev_2d(float *dx, size_t rows, size_t cols, size_t ld, uint64_t seed, uint64_t offset,
                               uint64_t gld, float lo, float hi, void *stream);

/* ---- convolution callers of the GEMM (the producer of BASELINE config 4's big operand).
 * Layouts are the reference's: planar C x H x W image, weights ch x (ich*k*k) row-major with column index
 * c*k*k + ki*k + kj, column matrix (ich*k*k) x (Ho*Wo), output ch x (Ho*Wo); square kernel / pad / stride.
 *   im2col_cuda            replaces the OpenCL `im2col` kernel + ocl_im2col (sgemm_ocl1.h:81-119,255-270) and the CPU
 *                          im2col (sgemm_gl1.h:166-190); host pointers, blocking.
 *   convolution_cuda       replaces ocl_convolution (sgemm_ocl1.h:271-300): outputs = weights . im2col(inputs).
 *   convolution_cuda_LReLU replaces gl_convolution_LReLU (sgemm_gl1.h:192-218) / the disabled ocl_convolution_LReLU
 *                          (sgemm_ocl1.h:301-338): + bias[ch] and LeakyReLU(0.1), FUSED into the GEMM epilogue instead of
 *                          the reference's separate host loop (sgemm_gl1.h:210-217).
 *   *_dev                  device pointers, asynchronous; d_workspace holds ich*k*k*Ho*Wo floats; d_bias may be NULL;
 *                          slope = 1 means no activation.  Returns 0 on success.
 *   convolution_cuda_batched_dev   `nimg` images [nimg][ich][h][w] -> [nimg][ch][Ho*Wo] with shared weights (BASELINE config 4 is 64 images
 *                          of 128 x 56 x 56, 256 filters 3 x 3: the GEMM M=256, N=64*3136, K=1152 in the reference's orientation).
 * Fused path (implicit GEMM), strides 1..8: the column matrix is never built -- one image-sized pass makes a channels-last copy of
 * the input (stream-ordered scratch, k*k times smaller than the column matrix) from which K1 gathers its B tiles with 4-D TMA
 * boxes (zero padding = TMA out-of-bounds fill); d_workspace is not touched and may be NULL.  Automatic rule: taken when the padded work (output width and channels rounded up to 32) stays within 30 %
 * of the real work, ch >= 64 and there are >= 256 output pixels; sgemm_cuda_set_conv_fusion(0 never | 1 whenever possible |
 * -1 rule); sgemm_cuda_last_conv_fused() reports what the last convolution did. */
int  convolution_cuda_batched_dev(int mode, void *stream, const float *d_inputs, int nimg, int ich, int w, int h, const float *d_weights,
                                  int k, int pad, int stride, float *d_outputs, int ch, const float *d_bias, float slope, float *d_workspace);
void sgemm_cuda_set_conv_fusion(int mode);
int  sgemm_cuda_last_conv_fused(void);
void im2col_cuda(const float *im, int channels, int height, int width, int k, int pad, int stride, float *col);
int  im2col_cuda_dev(const float *d_im, int channels, int height, int width, int k, int pad, int stride, float *d_col, void *stream);
void convolution_cuda(const float *inputs, int ich, int w, int h, const float *weights, int k, int pad, int stride,
                      float *outputs, int ch);
void convolution_cuda_LReLU(const float *inputs, int ich, int w, int h, const float *weights, int k, int pad, int stride,
                            float *outputs, int ch, const float *bias);
int  convolution_cuda_dev(int mode, void *stream, const float *d_inputs, int ich, int w, int h, const float *d_weights, int k,
                          int pad, int stride, float *d_outputs, int ch, const float *d_bias, float slope, float *d_workspace);

/* ---- level-1 / level-2 companions, HBM-bound.
 *   saxpy_cuda   replaces saxpy_cpu (ugemm.h:75-86; also saxpy_avx ugemm.h:58-73 and the OpenCL Xaxpy kernel of
 *                saxpy_ocl.c:129-157, whose host code uploads x and y, runs, downloads y):  y[i*incy] += alpha * x[i*incx],
 *                one fused multiply-add per element.  Host pointers, blocking.  incx, incy >= 1.
 *   sgemv_cuda   replaces sgemv_cpu (ugemm.h:124-150), argument for argument:  y[m*incy] = alpha * sum_n a(m,n) * x[n*incx] + beta * y[m*incy]
 *                for m < M, n < N, where a(m,n) = A[m + n*lda] when trans == 'N' (lda >= M) and A[n + m*lda] otherwise
 *                ('T', lda >= N) -- M is the length of y and N the length of x in BOTH cases, as in the reference.
 *                Deliberate differences: x is strided by incx (the reference strides x by incy, ugemm.h:140,147 -- a typo
 *                that only shows when incx != incy); beta == 0 overwrites y without reading it (consistent with sgemm_cuda);
 *                trans letters other than N/T (either case) are an error instead of meaning 'T'.
 *   *_dev        device pointers, asynchronous on `stream`.  Return 0 on success. */
void saxpy_cuda(int N, float alpha, const float *x, int incx, float *y, int incy);
int  saxpy_cuda_dev(void *stream, int N, float alpha, const float *dx, int incx, float *dy, int incy);
void sgemv_cuda(char trans, int M, int N, float alpha, const float *A, int lda, const float *x, int incx,
                float beta, float *y, int incy);
int  sgemv_cuda_dev(void *stream, char trans, int M, int N, float alpha, const float *dA, int lda, const float *dx, int incx,
                    float beta, float *dy, int incy);

/* ---- DGEMM: the path of check_dgemm.c.  Same 14-argument signature in double as dgemm_cpu (ugemm.h:162-178), _dgemm_c
 * (gemm_cpu.h:284-298 instantiated with real = double, ugemm.h:29-33) and dgemm_avx (dgemm_avx.h:844-858), i.e. usable as the
 * `uut` of test_dgemm (check_dgemm.c:86-97,255-258).  One kernel (K4, FP64 tensor-core mma.sync inner loop, FP64-pipe bound); semantics,
 * quirk decisions and error behaviour are those of sgemm_cuda.  dgemm_cuda: host pointers, blocking.  dgemm_cuda_dev:
 * device pointers, asynchronous.  dgemm_cuda_time_dev: mean / best of `iters` launches by CUDA events.  0 on success. */
void dgemm_cuda(char major, char transA, char transB, int M, int N, int K, double alpha,
                const double *A, int lda, const double *B, int ldb, double beta, double *C, int ldc);
int  dgemm_cuda_dev(void *stream, char major, char transA, char transB, int M, int N, int K, double alpha,
                    const double *dA, int lda, const double *dB, int ldb, double beta, double *dC, int ldc);
int  dgemm_cuda_time_dev(int iters, int warmup, char major, char transA, char transB, int M, int N, int K, double alpha,
                         const double *dA, int lda, const double *dB, int ldb, double beta, double *dC, int ldc,
                         float *ms_avg, float *ms_min);

/* ---- hardware probe used by tests/DESIGN.md: runs one 128 x 16 x (8*ksteps) TF32 tcgen05 product
 * chain on raw fp32 bit patterns and returns the 128x16 fp32 accumulator, so the rounding behaviour of
 * the tensor core (operand truncation, accumulator rounding) can be pinned.  A: 128 x 8*ksteps row-major,
 * B: 16 x 8*ksteps row-major (K-major both), D: 128 x 16.  Host pointers.  Returns 0 on success. */
int ugemm_cuda_probe_tf32(const float *A, const float *B, float *D, int ksteps);

#ifdef __cplusplus
}
#endif
#endif /* UGEMM_CUDA_H */
