/*
 * kgb200.h — C ABI of libkgb200.so, the B200 (sm_100a) array-primitive library behind the
 * KlongPy backend seam.
 *
 * The reference (briangu/klongpy) has no FFI: every array primitive is a NumPy (or torch) call made from
 * Python.  Each entry point below therefore cites the *reference call site* whose array-library call it
 * replaces (paths relative to the reference root).  All pointers are raw device addresses unless a
 * parameter is documented as host memory.  Outputs and workspaces are allocated by the caller (two-phase
 * `*_workspace_bytes` query); the library never owns result memory and never synchronises the stream
 * unless a function is documented as doing so (kgb_argsort / kgb_group / kgb_unique_first / kgb_groupagg_f64 do).  Every function returns 0 on success and a negative
 * KGB_E* code on failure; kgb_last_error() returns a thread-local message.
 *
 * No torch types appear here.  `stream` is a CUstream / cudaStream_t passed as void*.
 */
#ifndef KGB200_H
#define KGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KGB_ABI_VERSION 1

/* error codes */
#define KGB_OK 0
#define KGB_EINVAL (-1)   /* bad argument / unsupported program                         */
#define KGB_ECUDA (-2)    /* CUDA runtime / driver error                               */
#define KGB_ECOMPILE (-3) /* NVRTC failed on a generated expression kernel             */
#define KGB_ENOGPU (-4)   /* no usable sm_100 device                                   */

/* element types (KlongPy arrays are int64 / float64; f32/i32/bool appear through interop) */
#define KGB_F64 0
#define KGB_I64 1
#define KGB_F32 2
#define KGB_I32 3
#define KGB_B8 4 /* numpy bool_ / torch.bool, one byte 0/1 */

/* argument kinds of an expression program */
#define KGB_ARG_ARRAY 0     /* device pointer to n contiguous elements                      */
#define KGB_ARG_SCALAR 1    /* HOST pointer to one 8-byte value (double or int64), by value */
#define KGB_ARG_DEVSCALAR 2 /* device pointer to one element (0-d array), read in-kernel    */

/* what an expression kernel does with the per-element value of the program */
#define KGB_MODE_MAP 0    /* out[i] = f(args[i])                  (dyads.py:6-21 etc.)      */
#define KGB_MODE_REDUCE 1 /* out[0] = op-fold over i of f(args[i]) (adverbs.py:210-246)     */
#define KGB_MODE_SCAN 2   /* out[i] = op-fold over j<=i            (adverbs.py:316-334)     */
#define KGB_MODE_SCANRED 3 /* out[0] = op2-fold_i g(scan_op(f)[i], args[i])   (|/(|\x)-x)   */

/* fold operators for REDUCE / SCAN (IR ops '+','*','|','&', base.py:328-336) */
#define KGB_RED_ADD 0
#define KGB_RED_MUL 1
#define KGB_RED_MAX 2
#define KGB_RED_MIN 3

/* postfix opcodes of an expression program (one int32 each; PUSH_* carry operands inline) */
enum kgb_opcode {
    KGB_OP_ARG = 1,     /* + 1 word: argument index                                         */
    KGB_OP_LIT_I64 = 2, /* + 2 words: low, high 32 bits of the int64                        */
    KGB_OP_LIT_F64 = 3, /* + 2 words: low, high 32 bits of the IEEE-754 double              */
    KGB_OP_SCANVAL = 4, /* (SCANRED epilogue only) the running scan value at element i      */
    /* dyads — compiler IR 'binop' (numpy_backend.py:126-134), 'cmp' (:136-145) and the eager
       dyads the IR lacks (dyads.py:649-700 max/min, :785-803 fmod) */
    KGB_OP_ADD = 16,
    KGB_OP_SUB = 17,
    KGB_OP_MUL = 18,
    KGB_OP_DIV = 19, /* true divide, ints -> f64 (dyads.py:214-232)                          */
    KGB_OP_POW = 20, /* numpy ** : x**2 == x*x, int**int stays int (numpy_backend.py:131)     */
    KGB_OP_EQ = 21,  /* -> i64 0/1 (compiled form ((l==r)*1), numpy_backend.py:145)           */
    KGB_OP_GT = 22,
    KGB_OP_LT = 23,
    KGB_OP_MAX = 24, /* NaN-propagating like np.maximum (dyads.py:673)                        */
    KGB_OP_MIN = 25,
    KGB_OP_FMOD = 26, /* C truncated remainder like np.fmod (dyads.py:803)                    */
    KGB_OP_EQ_B = 27, /* raw comparisons -> bool, the np.less/np.greater/== the eager verbs   */
    KGB_OP_GT_B = 28, /*   call before kg_truth (dyads.py:609,722; base.py:176-182)           */
    KGB_OP_LT_B = 29,
    KGB_OP_NE_B = 30,
    KGB_OP_GE_B = 31,
    KGB_OP_LE_B = 32,
    KGB_OP_ISCLOSE = 33, /* |a-b| <= atol + rtol*|b| with numpy defaults (adverbs.py:33)      */
    KGB_OP_FLOORDIV = 34, /* np.floor_divide: integers stay integers, exact (x//0 -> 0); floats follow
                             numpy's divmod (the array type's `//`)                              */
    KGB_OP_MOD = 35,     /* np.remainder: floored, sign of the divisor (the array type's `%`)    */
    /* monads */
    KGB_OP_NEG = 48,      /* monads.py:224-237                                                */
    KGB_OP_ABS = 49,      /* monads.py:408-427                                                */
    KGB_OP_FLOOR = 50,    /* np.floor, float result                                           */
    KGB_OP_FLOOR_I64 = 51, /* floor then astype(int)  (base.py:198-203)                        */
    KGB_OP_TRUNC = 52,    /* dyads.py:754                                                     */
    KGB_OP_RECIP = 53,    /* np.reciprocal(asarray(x,float)) (monads.py:316)                  */
    KGB_OP_NOT_B = 54,    /* np.logical_not -> bool (monads.py:256)                           */
    KGB_OP_TO_F64 = 55,
    KGB_OP_TO_I64 = 56,   /* astype(int): C truncation (base.py:192-196)                      */
    KGB_OP_TO_F32 = 57,
    KGB_OP_TO_B8 = 58,
    KGB_OP_TO_I32 = 59,
    KGB_OP_SQRT = 64, /* the .bkf("name") unary math pulled from backend.np (sys_fn.py:715-723) */
    KGB_OP_EXP = 65,
    KGB_OP_LOG = 66,
    KGB_OP_SIN = 67,
    KGB_OP_COS = 68,
    KGB_OP_TAN = 69,
    KGB_OP_TANH = 70,
    KGB_OP_CEIL = 71,
    KGB_OP_SIGN = 72,
    KGB_OP_ISNAN_B = 73,
    KGB_OP_ISINF_B = 74,
    KGB_OP_LOG2 = 75,
    KGB_OP_LOG10 = 76,
    KGB_OP_SINH = 77,
    KGB_OP_COSH = 78,
    KGB_OP_ASIN = 79,
    KGB_OP_ACOS = 80,
    KGB_OP_ATAN = 81,
    KGB_OP_EXPM1 = 82,
    KGB_OP_LOG1P = 83,
    KGB_OP_SQUARE = 84,
    KGB_OP_WHERE = 96, /* ternary: pops (cond, a, b) -> cond ? a : b                          */
};

#define KGB_MAX_ARGS 12

typedef struct kgb_expr kgb_expr; /* opaque compiled expression kernel set */

/* ---------------------------------------------------------------- library / device -------------- */
int kgb_abi_version(void);
const char* kgb_last_error(void);
/* Bind the calling thread to `device` (primary context) and verify it is compute capability 10.x.
   Must be called once per thread before any other call (klongpy backends are per-interpreter,
   interpreter.py:249). */
int kgb_init(int device);
int kgb_device_sm_count(int device, int* out);

/* ---------------------------------------------------------------- fused expression kernels ------ */
/* Replaces NumpyBackendProvider.compile_expr_ir / _ir_to_source (numpy_backend.py:105-178): one
   single-pass kernel per IR tree instead of one NumPy call (and temporary) per node.  Also the engine
   behind every eager dyad/monad (one-node programs).
   code/n_code: postfix program.  For KGB_MODE_REDUCE an optional code2 is a SECOND expression over the same
   arguments folded with red_op2 in the same pass (both accumulators 8 bytes wide): out[0] receives the first
   result, the next 8-byte slot the second — `(+/p*s)%+/s` (examples/bench_compiler.kg VWAP) reads p and s once.
   For KGB_MODE_SCANRED, code2/n_code2 is the epilogue program evaluated
   at each element with KGB_OP_SCANVAL available, folded with red_op2.  For KGB_MODE_SCAN an optional
   code2 over scalar arguments only is the SEED folded in front of element 0 (the exclusive offset a shard
   carries in a multi-GPU scan).  Otherwise pass NULL/0.
   out_dtype receives the result element type. */
int kgb_expr_compile(const int32_t* code, int32_t n_code, const int32_t* code2, int32_t n_code2,
                     const int32_t* arg_kinds, const int32_t* arg_dtypes, int32_t n_args, int32_t mode,
                     int32_t red_op, int32_t red_op2, kgb_expr** out_handle, int32_t* out_dtype);
/* kgb_expr_compile for a two-output REDUCE (code + code2) with an EPILOGUE: `code3` is a program over two arguments —
   argument 0 the finished first fold, argument 1 the second — evaluated once by the kernel's last CTA; the launch then
   writes ONE value of type *out_dtype to `out` (`(+/p*s)%+/s`, compiler.py's VWAP example: both sums and the divide
   in one launch instead of a fold kernel plus a one-thread kernel).  kgb_expr_out_dtype2 returns -1 for such handles. */
int kgb_expr_compile_epi(const int32_t* code, int32_t n_code, const int32_t* code2, int32_t n_code2, const int32_t* code3,
                         int32_t n_code3, const int32_t* arg_kinds, const int32_t* arg_dtypes, int32_t n_args,
                         int32_t mode, int32_t red_op, int32_t red_op2, kgb_expr** out_handle, int32_t* out_dtype);
/* Same code generation + NVRTC compile to an sm_100a cubin, but WITHOUT loading it: works on a box with
   no GPU (build check, host-logic tests).  *source_out stays valid until the next call on the thread. */
int kgb_expr_check(const int32_t* code, int32_t n_code, const int32_t* code2, int32_t n_code2,
                   const int32_t* arg_kinds, const int32_t* arg_dtypes, int32_t n_args, int32_t mode,
                   int32_t red_op, int32_t red_op2, int32_t* out_dtype, const char** source_out);
/* bytes of zero-initialisable scratch a launch over n elements needs (0 for MAP) */
int kgb_expr_workspace_bytes(const kgb_expr* h, int64_t n, size_t* out_bytes);
/* args[i]: device pointer (ARRAY, DEVSCALAR) or host pointer to 8 bytes (SCALAR).  out: device.
   REDUCE over n == 0 is rejected (the host mirrors numpy's identity / ValueError rules). */
int kgb_expr_launch(kgb_expr* h, const void* const* args, int64_t n, void* out, void* workspace,
                    size_t workspace_bytes, void* stream);
/* kgb_expr_launch with flags.  KGB_LAUNCH_WS_CLEAN: the caller zero-filled `workspace` when it allocated it and has
   since passed it only to kgb_expr_launch* calls on this stream — then a REDUCE launch needs no memset in front of it
   (the fold's last CTA returns its ticket word to zero; SCAN launches zero what they use themselves either way).
   Saves one driver call per fold: ~2 us of the ~10 us a small-vector `+/a` costs end to end. */
#define KGB_LAUNCH_WS_CLEAN 1
int kgb_expr_launch_ex(kgb_expr* h, const void* const* args, int64_t n, void* out, void* workspace,
                       size_t workspace_bytes, void* stream, int32_t flags);
/* Host-only type inference: the element type a MAP program would produce for these argument kinds / dtypes (the
   NumPy promotion rules of numpy_backend.py's generated source).  No NVRTC, no GPU: lets the host answer `.dtype`
   for a deferred (not yet launched) elementwise expression. */
int kgb_expr_infer(const int32_t* code, int32_t n_code, const int32_t* arg_kinds, const int32_t* arg_dtypes,
                   int32_t n_args, int32_t* out_dtype);
/* element type of the second result of a two-output reduce, -1 if the handle has none */
int kgb_expr_out_dtype2(const kgb_expr* h);
/* generated CUDA source of a compiled expression (for tests, DESIGN.md and ncu --import-source) */
const char* kgb_expr_source(const kgb_expr* h);
/* name of the dominant kernel in the compiled module (shows up in ncu) */
const char* kgb_expr_kernel_name(const kgb_expr* h);

/* ---------------------------------------------------------------- grade (argsort) --------------- */
/* Replaces backend.argsort -> np.argsort (numpy_backend.py:75-80) behind kg_argsort's numeric fast
   path (writer.py:114-117).  Stable ascending LSD radix sort of (key, position); `descending` returns
   the stable ascending permutation reversed, which is the language contract (writer.py:97-108,
   tests/test_suite.py:82).  idx_out: int64[n].  keys_sorted_out: optional (NULL) dtype-typed [n].
   SYNCHRONISES `stream` for n >= 2^21: a 64 K-key sample is read back once to pick the digit plan
   (live key bits and position packed into one 64-bit word, 7..10-bit digits; csrc/kgb_sort_packed.cuh),
   and keys wider than the packed word are read back once more after the prefix fix-up.  kgb_group,
   kgb_unique_first and everything else built on the sort inherit this. */
int kgb_argsort_workspace_bytes(int32_t dtype, int64_t n, size_t* out_bytes);
int kgb_argsort(const void* keys, int32_t dtype, int64_t n, int32_t descending, int64_t* idx_out,
                void* keys_sorted_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- group / find / range ---------- */
/* Group `=a` (monads.py:182-204: np.unique(return_inverse) + one np.where per group, O(G*N)):
   order_out int64[n] = stable argsort; starts_out int64[n+1] receives the G group start offsets into
   order_out followed by n; ngroups_out: device int64[1].  Groups are in ascending key order, indices
   ascending within a group. */
int kgb_group_workspace_bytes(int32_t dtype, int64_t n, size_t* out_bytes);
int kgb_group(const void* keys, int32_t dtype, int64_t n, int64_t* order_out, int64_t* starts_out,
              int64_t* ngroups_out, void* workspace, size_t workspace_bytes, void* stream);
/* Find `a?b`, atom b (dyads.py:342: np.where(np.asarray(a)==b)[0]) and Where on 0/1 flags: writes the
   ascending positions of nonzero flags (or of elements equal to *needle, a HOST 8-byte value of `dtype`)
   to pos_out int64[n]; count_out: device int64[1]. */
int kgb_select_workspace_bytes(int64_t n, size_t* out_bytes);
int kgb_find_equal(const void* a, int32_t dtype, int64_t n, const void* needle_host, int64_t* pos_out,
                   int64_t* count_out, void* workspace, size_t workspace_bytes, void* stream);
int kgb_nonzero(const void* flags, int32_t dtype, int64_t n, int64_t* pos_out, int64_t* count_out,
                void* workspace, size_t workspace_bytes, void* stream);
/* Range `?a` (monads.py:260-295): unique elements in order of first appearance.  out: dtype-typed [n];
   first_idx_out int64[n] (positions of the first occurrences, ascending); count_out device int64[1]. */
int kgb_unique_first_workspace_bytes(int32_t dtype, int64_t n, size_t* out_bytes);
int kgb_unique_first(const void* a, int32_t dtype, int64_t n, void* out, int64_t* first_idx_out,
                     int64_t* count_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- support verbs ----------------- */
/* !n (monads.py:52: np.arange) */
int kgb_iota_i64(int64_t* out, int64_t n, int64_t start, int64_t step, void* stream);
/* a@b with integer list b (dyads.py:176-181: Python loop a[x] for x in b); negative indices wrap like
   numpy; out-of-range sets *err_flag (device int32) and writes 0. elem_bytes in {1,4,8}. */
int kgb_gather(const void* a, int32_t elem_bytes, int64_t n_a, const int64_t* idx, int64_t n_idx,
               void* out, int32_t* err_flag, void* stream);
/* &a (monads.py:79: np.repeat(np.arange(len(a)), a)): counts int64[n] >= 0, offsets = exclusive scan of
   counts (caller computes with a SCAN program), total = sum.  out int64[total]. */
int kgb_expand(const int64_t* counts, const int64_t* incl_offsets, int64_t n, int64_t total,
               int64_t* out, void* stream);
/* |a reverse (monads.py:319-340) */
int kgb_reverse(const void* a, int32_t elem_bytes, int64_t n, void* out, void* stream);
/* backend.array_equal (base.py:222-224): *out_flag (device int32) = 1 iff all elements compare equal
   (NaN != NaN like np.array_equal). */
int kgb_array_equal(const void* a, const void* b, int32_t dtype, int64_t n, int32_t* out_flag,
                    void* stream);

/* ---------------------------------------------------------------- grouped aggregation ----------- */
/* 1brc-style grouped min/max/sum/count (examples/1brc/worker.kg:3-5 collect/merge/stats): keys int64 in
   [0, n_groups), vals f64.  Outputs f64 min[G], max[G], sum[G] and int64 count[G]; groups with no rows
   get +inf/-inf/0/0.  `accumulate` != 0 merges into existing tables (the `merge` step). */
int kgb_groupagg_workspace_bytes(int64_t n, int64_t n_groups, size_t* out_bytes);
int kgb_groupagg_f64(const int64_t* keys, const double* vals, int64_t n, int64_t n_groups,
                     double* min_out, double* max_out, double* sum_out, int64_t* count_out,
                     int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream);
/* The same without the synchronising read-back of the "key outside [0, n_groups)" flag: nothing blocks the host, so a
   caller that follows the aggregation with an exchange (dist.sharded_groupagg: NCCL collectives + fold) enqueues it
   while the aggregation kernel is still running.  kgb_groupagg_status(workspace, n_groups, stream) collects the flag
   later (one 4-byte read, synchronises `stream`; KGB_EINVAL if a key was out of range); `workspace` must stay
   untouched until then. */
int kgb_groupagg_f64_async(const int64_t* keys, const double* vals, int64_t n, int64_t n_groups, double* min_out,
                           double* max_out, double* sum_out, int64_t* count_out, int32_t accumulate, void* workspace,
                           size_t workspace_bytes, void* stream);
int kgb_groupagg_status(const void* workspace, int64_t n_groups, void* stream);
/* merge step of the multi-GPU exchange: dst tables (op)= src tables, elementwise over G entries */
int kgb_groupagg_merge(double* min_io, double* max_io, double* sum_io, int64_t* count_io,
                       const double* min_in, const double* max_in, const double* sum_in,
                       const int64_t* count_in, int64_t n_groups, void* stream);

/* The exchange of the sharded aggregation (north_star: "hash-partitioned NCCL all-to-all").  A rank's
   tables travel as one block of 8-byte lanes [4][G] = {min, max, sum as float64 bits, count as int64}.
   kgb_groupagg_fold: `n_blocks` blocks (block r at blocks + r*block_stride lanes), folded in block order -> tables;
     after one all_gather of the ranks' blocks this is the whole merge (small G), after the all-to-all it merges the
     owner's slices.
   kgb_groupagg_to_owners: local block -> send buffer [world][4][B], B = ceil(G/world); station s goes to its owner
     hash(s) = s mod world, slot s / world (empty slots carry +inf / -inf / 0 / 0).
   kgb_groupagg_from_owners: the all-gathered owner slices [world][4][B] -> the four tables in station order. */
int kgb_groupagg_fold(const void* blocks, int64_t n_blocks, int64_t n_groups, int64_t block_stride,
                      double* min_out, double* max_out, double* sum_out, int64_t* count_out, void* stream);
int kgb_groupagg_to_owners(const void* block, int64_t n_groups, int64_t world, void* send, void* stream);
int kgb_groupagg_from_owners(const void* full, int64_t n_groups, int64_t world, double* min_out, double* max_out,
                             double* sum_out, int64_t* count_out, void* stream);

/* ---------------------------------------------------------------- 1brc text ingest --------------- */
/* Replaces the per-line Python loop of examples/1brc/par.py:57-88 (`location, measurement = line.split(";")`,
   `np.asarray(temps, dtype=float)`) and the string grouping `=s` of worker.kg:5: a device buffer of complete lines
   `name;[-]d[d].d\n` becomes a dense station id and a float64 temperature per row plus the station dictionary.
   ONE pass over the text (round 2; round 1 read it twice):
     kgb_brc_parse  per 65536-byte tile: newline positions in shared memory, the tile's first row number from a
                    decoupled look-back over the tiles' line counts; per row: name span, multilinear hash, insert /
                    look up in an open-addressing table (names are compared byte for byte on a tag match),
                    fixed-point parse of the measurement (exactly float("[-]d[d].d")).  The text must end with
                    '\n'.  Rows r < rows_capacity are written: id_out[r] in [0, *n_names) in first-arrival order,
                    temp_out[r]; *rows_out_host receives the number of rows FOUND — when it exceeds rows_capacity
                    the outputs are incomplete and the call is repeated with larger arrays (rows_capacity 0 with
                    null outputs just counts and builds the dictionary).  Row i of dict_names_out (dict_capacity
                    rows of 256 bytes, 4-byte aligned) holds the dict_len_out[i] <= 255 bytes of name i,
                    zero-filled up to the next multiple of 16.  *n_names_out_host receives the dictionary size.
                    Synchronises the stream once (40-byte read-back).  Returns KGB_EINVAL for a malformed row, a
                    name longer than 255 bytes or more than dict_capacity distinct names.
   Because the row count sizes the outputs, two ways to get it first:
     kgb_brc_estimate  newline count of three 4 MB windows scaled to the text, + 2 % + 4096 rows (exact for texts up
                    to 12 MB): an upper estimate that is almost always enough, for the price of a 12 MB read;
     kgb_brc_scan   the exact count (one more read of the text).
   Both synchronise the stream once (8-byte read-back). */
int kgb_brc_workspace_bytes(int64_t n_bytes, size_t* out_bytes);
int kgb_brc_scan(const uint8_t* text, int64_t n_bytes, int64_t* rows_out_host, void* workspace,
                 size_t workspace_bytes, void* stream);
int kgb_brc_estimate(const uint8_t* text, int64_t n_bytes, int64_t* rows_estimate_out_host, void* workspace,
                     size_t workspace_bytes, void* stream);
int kgb_brc_parse(const uint8_t* text, int64_t n_bytes, int64_t rows_capacity, int64_t* id_out, double* temp_out,
                  uint8_t* dict_names_out, int32_t* dict_len_out, int64_t dict_capacity, int64_t* rows_out_host,
                  int64_t* n_names_out_host, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KGB200_H */
