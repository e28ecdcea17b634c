/*
 * tncb.h -- C ABI of libtncb200: the B200-native pairwise tensor-contraction hot
 * path of qc-tum/TNC (tnc v1.0.0 @ 5dd62b3).
 *
 * This is the drop-in boundary.  TNC reaches its numeric kernel through plain
 * Rust calls into the un-vendored crate `tetra` (tnc/src/tensornetwork/
 * contraction.rs:3,78-84); a maintainer replaces those calls by the entry
 * points below through an `extern "C"` block (see INTEGRATION.md for the Rust
 * binding).  Every entry point cites the reference interface it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions (identical to the reference):
 *   - elements are complex128, passed as interleaved (re, im) doubles;
 *   - data is row-major (C order) over the tensor's leg order;
 *   - the result of contracting a and b has legs (b \ a) ++ (a \ b)
 *     (tnc/src/tensornetwork/tensor.rs:463-479 via contraction.rs:64);
 *   - paths are "replace-left": (i, j) stores the result in slot i
 *     (tnc/src/contractionpath.rs:29-35).
 *
 * Errors: every function returns 0 (TNCB_OK) or a negative tncb_status; nothing
 * aborts.  The reference panics instead (tensordata.rs:42, contraction.rs:50);
 * the language shim maps non-zero to panic!/exception.
 * Threading: a tncb_ctx is single-threaded like the reference's driver loop;
 * distinct contexts (devices) may be driven from distinct threads.
 * There is NO CPU fallback: without a CUDA device tncb_ctx_create fails with
 * TNCB_ERR_CUDA.
 */
#ifndef TNCB_H
#define TNCB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tncb_status {
  TNCB_OK = 0,
  TNCB_ERR_INVALID = -1,      /* bad argument / malformed network or path            */
  TNCB_ERR_SHAPE = -2,        /* bond dimensions of a shared leg disagree            */
  TNCB_ERR_UNCONTRACTED = -3, /* slot already consumed / no data (tensordata.rs:42)  */
  TNCB_ERR_NOT_CONTRACTED = -4, /* >1 tensor left ("Not fully contracted", contraction.rs:50) */
  TNCB_ERR_OOM = -5,          /* device arena exhausted (peak live bytes > capacity) */
  TNCB_ERR_CUDA = -6,         /* CUDA runtime error (tncb_last_error has the text)   */
  TNCB_ERR_GATE = -7,         /* unknown gate / wrong angle count (gates.rs:54,103)  */
  TNCB_ERR_NCCL = -8,         /* NCCL error or libnccl not loadable                  */
  TNCB_ERR_UNSUPPORTED = -9   /* e.g. TensorData::File payloads (no HDF5)            */
} tncb_status;

typedef struct tncb_ctx tncb_ctx;       /* one device + stream + arena              */
typedef struct tncb_tensor tncb_tensor; /* a device-resident complex128 tensor      */
typedef struct tncb_plan tncb_plan;     /* a compiled (network, path) schedule      */

const char* tncb_strerror(int status);
/* Text of the last error raised on this thread (CUDA/NCCL message, offending pair). */
const char* tncb_last_error(void);
/* Library version / build info (arch list) -- cheap, needs no GPU. */
const char* tncb_version(void);

/* ---- context ------------------------------------------------------------------ */
/* arena_bytes = 0: grow on demand up to the device's free memory. */
int tncb_ctx_create(int device, size_t arena_bytes, tncb_ctx** out);
void tncb_ctx_destroy(tncb_ctx* ctx);
int tncb_ctx_synchronize(tncb_ctx* ctx);
/* The CUDA stream every kernel of this ctx is enqueued on (a cudaStream_t). */
void* tncb_ctx_stream(tncb_ctx* ctx);
/* Counters since creation / last reset: kernels launched by this library, arena peak. */
int tncb_ctx_stats(tncb_ctx* ctx, uint64_t* kernel_launches, uint64_t* arena_peak_bytes,
                   uint64_t* arena_live_bytes);
int tncb_ctx_reset_stats(tncb_ctx* ctx);
/* Dense-GEMM engine for large pairs (K >= 1536 and >= 96 output tiles of 128x128):
 * slices in [2,8] -> tcgen05 int8 path (exact 7-bit digit slicing; the default 8 slices cover the
 * full 53-bit mantissa, measured |err| ~ 1e-15..7e-15 of max|C|, like FP64 accumulation itself);
 * slices = 0 -> FP64 tensor pipe (DMMA) for every pair.  Smaller pairs always use DMMA / K0.
 * The environment variable TNCB_OZAKI_SLICES overrides the default. */
int tncb_ctx_set_tcgen05_slices(tncb_ctx* ctx, int slices);
/* Size thresholds of the tcgen05 engine (defaults: >= 96 output tiles of 128x128 and K >= 1536);
 * (1, 256) routes every pair with M, N, K >= 256 to it (used by the parity tests). */
int tncb_ctx_set_tcgen05_threshold(tncb_ctx* ctx, long long min_tiles, long long min_k);
/* Measurement aid: bracket the dominant GEMM kernel of every large pair (k1_kernel / oz_gemm_kernel)
 * with CUDA events on the ctx stream; tncb_ctx_last_gemm_ms synchronises and returns the last one. */
int tncb_ctx_time_gemm(tncb_ctx* ctx, int enable);
int tncb_ctx_last_gemm_ms(tncb_ctx* ctx, float* ms);

/* ---- tensors: replaces tetra::Tensor::{new_from_flat, elements, shape, ndim}
 *      (tnc/src/tensornetwork/tensordata.rs:31-37, tnc/src/io/hdf5.rs:105-106) ---- */
int tncb_tensor_upload(tncb_ctx* ctx, int rank, const uint64_t* dims,
                       const double* host_re_im, tncb_tensor** out);
int tncb_tensor_alloc(tncb_ctx* ctx, int rank, const uint64_t* dims, tncb_tensor** out);
int tncb_tensor_download(tncb_ctx* ctx, const tncb_tensor* t, double* host_re_im);
/* Asynchronous variants on the ctx stream (host buffer should be pinned; the caller
 * synchronises with tncb_ctx_synchronize before touching it). */
int tncb_tensor_write(tncb_ctx* ctx, tncb_tensor* t, const double* host_re_im);
int tncb_tensor_read(tncb_ctx* ctx, const tncb_tensor* t, double* host_re_im);
int tncb_tensor_free(tncb_ctx* ctx, tncb_tensor* t);
int tncb_tensor_rank(const tncb_tensor* t);
int tncb_tensor_dims(const tncb_tensor* t, uint64_t* dims_out);
uint64_t tncb_tensor_elements(const tncb_tensor* t);
void* tncb_tensor_device_ptr(const tncb_tensor* t); /* double2*, row-major */

/* ---- one pairwise contraction: replaces
 *      tetra::contract(out_legs, a_legs, a, b_legs, b) as called at
 *      tnc/src/tensornetwork/contraction.rs:78-84.  Consumes a and b (the Rust
 *      call moves them), returns a new tensor whose legs are out_legs, which must
 *      equal (b \ a) ++ (a \ b).  Pass out_legs = NULL to skip that check and
 *      read the legs back with tncb_pair_out_legs. ---- */
int tncb_contract_pair(tncb_ctx* ctx, int n_out, const uint64_t* out_legs,
                       int n_a, const uint64_t* a_legs, tncb_tensor* a,
                       int n_b, const uint64_t* b_legs, tncb_tensor* b,
                       tncb_tensor** out);
/* Same, but a and b stay alive (for benchmarking a single pair repeatedly). */
int tncb_contract_pair_keep(tncb_ctx* ctx, int n_a, const uint64_t* a_legs, const tncb_tensor* a,
                            int n_b, const uint64_t* b_legs, const tncb_tensor* b,
                            tncb_tensor** out);
/* Into a caller-provided output tensor (no allocation inside the timed region). */
int tncb_contract_pair_into(tncb_ctx* ctx, int n_a, const uint64_t* a_legs, const tncb_tensor* a,
                            int n_b, const uint64_t* b_legs, const tncb_tensor* b,
                            tncb_tensor* out);
/* Leg algebra only (no GPU): Tensor::symmetric_difference, tensor.rs:463-479.
 * Writes (b\a)++(a\b) and the GEMM view M=|a\b|, N=|b\a|, K=|a&b|. */
int tncb_pair_out_legs(int n_a, const uint64_t* a_legs, const uint64_t* a_dims,
                       int n_b, const uint64_t* b_legs, const uint64_t* b_dims,
                       int* n_out, uint64_t* out_legs, uint64_t* out_dims,
                       uint64_t* m, uint64_t* n, uint64_t* k);
/* Which kernel class the planner would pick for that pair (no GPU):
 * 0 = K0 strided/warp-reduce kernel, 1 = K1 fused gather ZGEMM (DMMA, or tcgen05 K1' above the size
 * threshold), 2 = K2 streaming kernel (big tensor x tiny tensor, HBM-bound). */
int tncb_pair_kernel_class(int n_a, const uint64_t* a_legs, const uint64_t* a_dims,
                           int n_b, const uint64_t* b_legs, const uint64_t* b_dims);

/* ---- tetra::Tensor::transpose / conjugate
 *      (tnc/src/builders/circuit_builder.rs:106 Permutor::apply; gates.rs:87,98) ---- */
/* out dims[i] = in dims[perm[i]] (numpy.transpose semantics); consumes `t`. */
int tncb_permute(tncb_ctx* ctx, tncb_tensor* t, const int* perm, tncb_tensor** out);
int tncb_conjugate(tncb_ctx* ctx, tncb_tensor* t); /* in place */
/* dst += src (same element count): accumulation of sliced contractions (the reference's declared
 * future work, book/src/future_work.md:9-11). */
int tncb_tensor_add(tncb_ctx* ctx, tncb_tensor* dst, const tncb_tensor* src);

/* ---- gate table: replaces load_gate / load_gate_adjoint (tnc/src/gates.rs:50-66).
 *      Host-side; writes 4 or 16 interleaved complex values, *rank = 2 or 4. ---- */
int tncb_gate_matrix(const char* name, const double* angles, int n_angles, int adjoint,
                     double* out_re_im, int* rank);

/* ---- networks: mirrors tnc::tensornetwork::tensor::Tensor (tensor.rs:21-37) and
 *      tnc::contractionpath::ContractionPath (contractionpath.rs:29-35) as plain
 *      C trees so that cgo / Rust repr(C) / ctypes can build them. ---- */
typedef enum tncb_data_kind {
  TNCB_DATA_UNCONTRACTED = 0, /* TensorData::Uncontracted (composite or empty slot) */
  TNCB_DATA_MATRIX = 1,       /* TensorData::Matrix: host_re_im, row-major          */
  TNCB_DATA_GATE = 2,         /* TensorData::Gate((name, angles, adjoint))          */
  TNCB_DATA_DEVICE = 3        /* already on the device (consumed by the call)       */
} tncb_data_kind;

typedef struct tncb_tn {
  /* composite: n_children > 0, legs ignored.  leaf: n_children == 0. */
  size_t n_children;
  const struct tncb_tn* children;
  int rank;
  const uint64_t* legs;
  const uint64_t* dims;
  int kind; /* tncb_data_kind */
  const double* host_re_im;
  const char* gate_name;
  const double* gate_angles;
  int n_gate_angles;
  int gate_adjoint;
  tncb_tensor* device;
} tncb_tn;

typedef struct tncb_path {
  size_t n_pairs;
  const uint64_t* pairs; /* i0, j0, i1, j1, ... replace-left */
  size_t n_nested;
  const uint64_t* nested_index; /* child indices that have their own path */
  const struct tncb_path* nested;
} tncb_path;

/* contract_tensor_network(tn, path) (tnc/src/tensornetwork/contraction.rs:30-52):
 * nested paths first (ascending child index), then the top-level pairs in order.
 * Leaves are materialised (gates.rs tables) and uploaded in ONE host->device
 * copy, every pair runs on the device without host round trips, and the result
 * stays on the device.  out_legs must have room for *n_out legs (<= 64). */
int tncb_contract_tensor_network(tncb_ctx* ctx, const tncb_tn* tn, const tncb_path* path,
                                 tncb_tensor** out, int* n_out, uint64_t* out_legs);

/* Compile once / execute many: the same circuit with different payloads
 * (e.g. other bitstrings or angles) re-uses the schedule, arena layout and the
 * captured CUDA graph. */
int tncb_plan_create(tncb_ctx* ctx, const tncb_tn* tn, const tncb_path* path, tncb_plan** out);
int tncb_plan_execute(tncb_ctx* ctx, tncb_plan* plan, const tncb_tn* tn,
                      tncb_tensor** out, int* n_out, uint64_t* out_legs);
/* Schedule facts: #pairs, sum 8MNK, sum 16(MK+KN+MN), peak arena bytes, #kernels. */
int tncb_plan_info(const tncb_plan* plan, uint64_t* n_pairs, double* flops, double* bytes,
                   uint64_t* peak_bytes, uint64_t* n_kernels);
void tncb_plan_destroy(tncb_plan* plan);

/* ---- partitioned fan-in: replaces tnc::mpi::communication
 *      (scatter_tensor_network :125-195, intermediate_reduce_tensor_network :199-249).
 *      One process per GPU; boundary tensors move GPU->GPU with ncclSend/ncclRecv
 *      on the ctx stream (no serialisation: legs/dims are derived on every rank). ---- */
/* 128-byte NCCL unique id; create on rank 0, broadcast out of band. */
int tncb_comm_unique_id(uint8_t id_out[128]);
int tncb_comm_init(tncb_ctx* ctx, int world_size, int rank, const uint8_t id[128]);
int tncb_comm_send(tncb_ctx* ctx, const tncb_tensor* t, int peer);
int tncb_comm_recv(tncb_ctx* ctx, int rank_dims, const uint64_t* dims, int peer, tncb_tensor** out);
/* In-place sum over all ranks (ncclAllReduce on the ctx stream): combines sliced contractions. */
int tncb_comm_allreduce_sum(tncb_ctx* ctx, tncb_tensor* t);
int tncb_comm_destroy(tncb_ctx* ctx);
/* get_tensor_mapping (mpi/communication.rs:89-115): partition -> rank, the
 * partition on the left of the last top-level pair goes to rank 0, the others
 * to 1, 2, ... in ascending partition index.  rank_of[p] for p < n_partitions. */
int tncb_fanin_mapping(size_t n_partitions, const uint64_t* partition_index,
                       size_t n_pairs, const uint64_t* toplevel_pairs, int world_size,
                       int* rank_of_partition);

#ifdef __cplusplus
}
#endif
#endif /* TNCB_H */
