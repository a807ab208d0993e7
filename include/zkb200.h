/* zkb200.h -- C ABI of libzkb200.so: the B200 (sm_100a) accelerator for the groth16::prove() hot path
 * of republicprotocol/zksnark-rs.
 *
 * The reference is a pure-Rust crate with NO FFI boundary: its extension seams are the generic
 * bounds of `prove<P,T,U,V>` (src/groth16/mod.rs:213-229).  This header is the boundary a Rust
 * `extern "C"` block (rust/zkb200_sys.rs) binds; each entry point cites the
 * reference code it replaces.
 *
 * Conventions
 *  - Field elements (Fr scalars, Fq coordinates) cross the ABI as 4 x uint64_t little-endian limbs
 *    of the CANONICAL residue (never Montgomery form).
 *  - G1 points: affine (x, y) = 8 x uint64_t.  G2 points: affine (x.c0, x.c1, y.c0, y.c1) =
 *    16 x uint64_t.  The identity is all-zero (it is a legal CRS entry, src/groth16/mod.rs:407).
 *  - Every function returns 0 (ZKB_OK) or a negative error code; nothing throws or aborts.  The
 *    Rust shim turns non-zero into panic!() to mirror the reference's panics
 *    (src/groth16/fr.rs:54, src/field/mod.rs:440).
 *  - Caller owns all host buffers.  The library owns device objects until the matching *_free.
 *  - Calls are synchronous on return.  A zkb_ctx is bound to one CUDA device and is not
 *    thread-safe; distinct contexts are independent.
 *  - There is NO CPU fallback: every entry point that computes fails with ZKB_ERR_CUDA when no
 *    sm_100-class device is present.
 */
#ifndef ZKB200_H
#define ZKB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKB_OK 0
#define ZKB_ERR_CUDA (-1)        /* CUDA runtime error / no device */
#define ZKB_ERR_ARG (-2)         /* invalid argument (null pointer, size mismatch, non power of two ...) */
#define ZKB_ERR_ALLOC (-3)       /* device or host allocation failed */
#define ZKB_ERR_UNSUPPORTED (-4) /* valid request this build does not implement */
#define ZKB_ERR_DIV_ZERO (-5)    /* the reference would panic: inverse of zero (fr.rs:54,69) */
#define ZKB_ERR_COMM (-6)        /* multi-GPU exchange: a peer did not arrive within the timeout */

typedef struct zkb_ctx zkb_ctx;
typedef struct zkb_qap zkb_qap;     /* device-resident QAP<CoefficientPoly<FrLocal>> (sparse evaluation rows) */
typedef struct zkb_crs zkb_crs;     /* device-resident (SigmaG1<G1Local>, SigmaG2<G2Local>) */
typedef struct zkb_bases zkb_bases; /* device-resident vector of G1 or G2 affine points */
typedef struct zkb_comm zkb_comm;   /* one rank's end of a multi-GPU communicator (exchange window in HBM, mapped by the peers) */

/* ---- context ------------------------------------------------------------------------------- */
int zkb_ctx_create(zkb_ctx** out, int device_id);
void zkb_ctx_destroy(zkb_ctx* ctx);
/* Message of the last error on this context ("" if none).  ctx may be NULL (-> global message). */
const char* zkb_last_error(const zkb_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t zkb_launch_count(const zkb_ctx* ctx);
/* Per-kernel-class timing with CUDA events on the launching stream.  zkb_profile(ctx, 1) clears and
 * starts recording, zkb_profile(ctx, 0) stops; zkb_profile_read sums the recorded launch durations
 * of one class: 1 NTT butterfly passes, 2 G1 bucket accumulation, 3 G2 bucket accumulation; *units is
 * the work those launches covered (NTT: elements x passes; accumulation: point records = points x
 * windows, an upper bound that counts the ~2^-c fraction of zero digits). */
int zkb_profile(zkb_ctx* ctx, int enable);
/* zkb_profile(ctx, 2) additionally brackets EVERY kernel launch with events (tracing; adds a few
 * microseconds per launch); zkb_trace_dump writes "stream,kernel,start_ms,end_ms,dur_ms" rows
 * (times relative to the zkb_profile call) -- the per-stage timeline the reference only prints as
 * averages (fr.rs:339-358). */
int zkb_trace_dump(zkb_ctx* ctx, const char* path);
int zkb_profile_read(zkb_ctx* ctx, int kind, double* total_ms, uint64_t* count, uint64_t* units);
/* Pinned host memory for buffers that are copied every proof (weights).  Optional. */
int zkb_host_alloc(void** out, size_t bytes);
void zkb_host_free(void* p);
/* Raw device memory (so a torch-free caller can keep inputs resident in HBM). */
int zkb_dev_alloc(zkb_ctx* ctx, void** out, size_t bytes);
void zkb_dev_free(zkb_ctx* ctx, void* p);
int zkb_memcpy_h2d(zkb_ctx* ctx, void* dst, const void* src, size_t bytes);
int zkb_memcpy_d2h(zkb_ctx* ctx, void* dst, const void* src, size_t bytes);
int zkb_sync(zkb_ctx* ctx);
/* The stream every kernel of this context is launched on (a cudaStream_t), for event timing. */
void* zkb_stream(zkb_ctx* ctx);

/* ---- QAP: replaces `QAP<CoefficientPoly<FrLocal>>` (groth16/mod.rs:60-67) and its constructor
 * `From<RootRepresentation>` (groth16/fr.rs:140-173).  The reference stores every u_i, v_i, w_i as
 * a dense coefficient vector (3*M*n field elements); this type stores the same polynomials by
 * their non-zero evaluations on the root domain -- exactly the `DummyRep` data model
 * (circuit/dummy_rep.rs:7-13).  Rows are CSR by wire: row i of u holds (gate index k, value) for
 * every root_k where u_i(root_k) != 0.  Fast domain (roots == NULL): the n-th roots of unity, gate
 * k <-> omega^k (omega = 5^((r-1)/n)); n must be a power of two, 2 <= n <= 2^27: O(n log n) NTTs.
 * Generic domain (roots != NULL): what `QAP::from(DummyRep)` accepts for parser-produced circuits
 * (fr.rs:140-173): same results as the reference's Lagrange interpolation (coefficient_poly.rs:159-190),
 * schoolbook product (:93-130) and long division (field/mod.rs:428-469), computed with dense O(n^2)
 * device kernels, n <= 32768 (the n x n table of Lagrange coefficients takes n^2 * 32 bytes of HBM: 2 GiB at 8192,
 * 32 GiB at 32768).  Larger parser circuits: re-index the gates onto the roots of unity (gate k <-> omega^k, padded to a
 * power of two with empty gates) -- a different but equivalent QAP of the same circuit, which the reference permits
 * (DummyRep.roots is a public field; CircuitInstance::new takes the root assignment as a closure, circuit/mod.rs:99-104). */
typedef struct {
  uint64_t n;         /* number of gates = qap.degree                                   */
  uint64_t m;         /* number of rows (wires incl. the unity wire) = qap.u.len()      */
  uint64_t n_input;   /* qap.input (number of verifier-visible wires, excl. unity)      */
  const uint64_t* row_ptr[3];  /* u, v, w: m+1 offsets into gate/coeff                  */
  const uint32_t* gate[3];     /* nnz gate indices (0-based)                            */
  const uint64_t* coeff[3];    /* nnz x 4 limbs, canonical                              */
  const uint64_t* roots;       /* NULL: the n-th roots of unity (n a power of two).  Else n x 4 limbs:
                                  explicit pairwise-distinct roots, gate k <-> roots[k] (the reference's
                                  ASTParser uses 1..=n, circuit/mod.rs:517); any n in [1, 32768]; dense
                                  O(n^2) interpolation / product / division on the device            */
} zkb_qap_host;
int zkb_qap_upload(zkb_ctx* ctx, const zkb_qap_host* qap, zkb_qap** out);
void zkb_qap_free(zkb_ctx* ctx, zkb_qap* qap);

/* ---- CRS: replaces `SigmaG1<G1Local>` / `SigmaG2<G2Local>` (groth16/mod.rs:105-121). ---------- */
typedef struct {
  uint64_t n;            /* xi1 / xi2 length = qap.degree; xi_t has n-1 entries (mod.rs:168) */
  uint64_t n_sum_gamma;  /* input+1 (mod.rs:154)  */
  uint64_t n_sum_delta;  /* m-input-1 (mod.rs:163) */
  const uint64_t* alpha1; const uint64_t* beta1; const uint64_t* delta1;  /* 8 limbs each */
  const uint64_t* xi1;        /* n x 8       */
  const uint64_t* xi_t;       /* (n-1) x 8   */
  const uint64_t* sum_gamma;  /* n_sum_gamma x 8 */
  const uint64_t* sum_delta;  /* n_sum_delta x 8 */
  const uint64_t* beta2; const uint64_t* gamma2; const uint64_t* delta2;  /* 16 limbs each */
  const uint64_t* xi2;        /* n x 16      */
} zkb_crs_host;
/* Upload a CRS computed elsewhere (e.g. by the reference's setup()).  rank/world shard the MSM
 * base vectors by contiguous index ranges across `world` processes (one GPU each); pass 0,1 for a
 * single GPU. */
int zkb_crs_upload(zkb_ctx* ctx, const zkb_crs_host* crs, int rank, int world, zkb_crs** out);
/* groth16::setup (groth16/mod.rs:134-197) on the device with the five secrets injected
 * (toxic = alpha, beta, gamma, delta, x; 5 x 4 limbs, all non-zero as random_elem guarantees,
 * fr.rs:90-99).  Base points are (G1::one()*69) and (G2::one()*96), fr.rs:106-113. */
int zkb_setup(zkb_ctx* ctx, const zkb_qap* qap, const uint64_t* toxic, int rank, int world, zkb_crs** out);
/* Copy the (full, unsharded; requires world == 1) CRS back in the zkb_crs_host layout.  The caller
 * provides every buffer of `dst` sized from dst->n / n_sum_gamma / n_sum_delta (query with
 * zkb_crs_dims first). */
int zkb_crs_dims(const zkb_crs* crs, uint64_t* n, uint64_t* n_sum_gamma, uint64_t* n_sum_delta);
int zkb_crs_download(zkb_ctx* ctx, const zkb_crs* crs, zkb_crs_host* dst);
void zkb_crs_free(zkb_ctx* ctx, zkb_crs* crs);

/* ---- the hot path: groth16::prove (groth16/mod.rs:213-296) --------------------------------- */
typedef struct {
  uint64_t a[8];   /* Proof.a: G1 affine */
  uint64_t b[16];  /* Proof.b: G2 affine */
  uint64_t c[8];   /* Proof.c: G1 affine */
} zkb_proof;
/* weights: m x 4 limbs (host memory; copied to the device inside the call).  r, s: the two
 * scalars the reference draws at mod.rs:231, injected so results are reproducible. */
int zkb_prove(zkb_ctx* ctx, const zkb_qap* qap, const zkb_crs* crs, const uint64_t* weights,
              const uint64_t r[4], const uint64_t s[4], zkb_proof* out);
/* Same with the weights already resident in device memory (m x 4 limbs, canonical). */
int zkb_prove_dev(zkb_ctx* ctx, const zkb_qap* qap, const zkb_crs* crs, const uint64_t* d_weights,
                  const uint64_t r[4], const uint64_t s[4], zkb_proof* out);
/* Throughput mode: `count` independent proofs over the same QAP and CRS (weights[i]: m x 4 limbs each,
 * all host or all device pointers; r, s: count x 4 limbs).  Several proofs are kept in flight (two at
 * 2^20 gates -- three when the weights come from host memory, whose copy a third proof hides --, three up to 2^19, four
 * up to 2^17) on as many internal stream sets so that the short /
 * low-occupancy stages of one proof (polynomial stage, record sort, bucket reduction) overlap the
 * SM-filling bucket accumulation of another.  Results
 * are identical to `count` calls of zkb_prove (over a sharded CRS: of zkb_prove_partial, `out` then holds
 * this rank's partial records).  Host weights should be pinned (zkb_host_alloc) for the
 * copies to overlap. */
int zkb_prove_batch(zkb_ctx* ctx, const zkb_qap* qap, const zkb_crs* crs, const uint64_t* const* weights,
                    int weights_on_device, const uint64_t* r, const uint64_t* s, size_t count, zkb_proof* out);
/* Multi-GPU: each rank runs the polynomial stage and the MSMs over ITS shard of the CRS (rank 0 also
 * carries the fixed-point terms alpha1 + r delta1 etc.) and returns its partial sums of A, B, C in the
 * zkb_proof layout (affine, canonical: a 8 | b 16 | c 8 = 32 limbs).  The caller all-gathers the
 * 32-limb records (NCCL; EC addition is not an NCCL reduction op) and every rank -- or rank 0 --
 * folds them with zkb_prove_combine. */
#define ZKB_PARTIAL_LIMBS 32
int zkb_prove_partial(zkb_ctx* ctx, const zkb_qap* qap, const zkb_crs* crs, const uint64_t* weights,
                      int weights_on_device, const uint64_t r[4], const uint64_t s[4],
                      uint64_t* out_partial /* 32 limbs, host */);
int zkb_prove_combine(zkb_ctx* ctx, const uint64_t* partials /* world x 32, host */, int world, zkb_proof* out);
/* The same for `count` proofs at once.  zkb_prove_batch over a SHARDED CRS returns this rank's `count`
 * partial records (count x 32 limbs); all-gathering them gives partials[world][count][32], which one
 * kernel launch folds into `count` proofs. */
int zkb_prove_combine_batch(zkb_ctx* ctx, const uint64_t* partials /* world x count x 32, host */, int world,
                            size_t count, zkb_proof* out);
/* ---- ONE proof over several GPUs, exchanges inside the library (replaces the single-threaded
 * groth16/mod.rs:213-296 as a whole -- the reference has no process or device boundary, so this surface is new).
 * One process (or one context) per GPU, world = 2^k ranks.  Every rank owns an exchange window in its HBM that the
 * peers map (CUDA IPC between processes; the raw pointer between contexts of one process) and the kernels write
 * into directly over NVLink: no NCCL call and no host bounce on the data path.
 *   1. zkb_comm_create on every rank -> a 128-byte handle; 2. the host side gathers the `world` handles in rank order
 *   (any transport: torch.distributed / MPI / a file -- once, at start-up); 3. zkb_comm_connect maps the peers.
 * The polynomial stage shards the NTT's outer dimension: rank r evaluates gates r, r+G, ..., each of the six
 * transforms is a local size-n/G transform plus ONE all-to-all plus log G butterfly stages (three exchanges per
 * proof: 3 + 2 + 1 vectors of n/G elements per rank), and u_sum, v_sum, h come out in the "strided block" layout
 * (local index k1*q + t <-> coefficient (r*q + t) + (n/G)*k1, q = n/G^2).  zkb_setup_shard / zkb_crs_upload_shard
 * shard sigma_g1.xi, sigma_g1.xi_t and sigma_g2.xi the same way (sum_delta by contiguous ranges; the fixed points on
 * rank 0), so every rank's MSMs run over 1/G of the points; the 256-byte partial sums of A, B, C are written into
 * every peer's window and folded on the device ("allreduce" = all-to-all stores + local fold: EC addition is not a
 * reduction operator of NCCL).  Every rank returns the complete proof, bit-identical to zkb_prove on one GPU.
 * Requirements: roots-of-unity domain, n >= 2 * world^2, world <= 16.  Collective semantics: every rank must make the
 * same sequence of zkb_prove_shard* / zkb_ntt_shard calls; a rank that never arrives makes the others fail with
 * ZKB_ERR_COMM after ZKB_COMM_TIMEOUT_MS (default 20000) instead of hanging. */
#define ZKB_COMM_HANDLE_BYTES 128
int zkb_comm_create(zkb_ctx* ctx, int rank, int world, uint32_t max_log_n, zkb_comm** out, uint8_t* handle /* 128 bytes */);
int zkb_comm_connect(zkb_comm* comm, const uint8_t* handles /* world x 128 bytes, rank order */);
void zkb_comm_destroy(zkb_comm* comm);
/* rank / world of the communicator and its sticky status word (0 ok, 1 = an exchange timed out); any may be NULL */
int zkb_comm_info(const zkb_comm* comm, int* rank, int* world, int* status);
/* groth16::setup (mod.rs:134-197) / upload of a reference-made CRS for this rank's shard in the layout above */
int zkb_setup_shard(zkb_ctx* ctx, const zkb_comm* comm, const zkb_qap* qap, const uint64_t* toxic, zkb_crs** out);
int zkb_crs_upload_shard(zkb_ctx* ctx, const zkb_comm* comm, const zkb_crs_host* crs, zkb_crs** out);
/* groth16::prove over all ranks.  weights: the full witness (m x 4 limbs) on every rank, host or device. */
int zkb_prove_shard(zkb_ctx* ctx, zkb_comm* comm, const zkb_qap* qap, const zkb_crs* crs, const uint64_t* weights,
                    int weights_on_device, const uint64_t r[4], const uint64_t s[4], zkb_proof* out);
/* The same split in two so that several proofs (lanes 0..3, one exchange channel each) can be in flight, or several
 * ranks can be driven from one host thread: enqueue returns as soon as the work is queued, collect waits for it. */
int zkb_prove_shard_enqueue(zkb_ctx* ctx, zkb_comm* comm, const zkb_qap* qap, const zkb_crs* crs, const uint64_t* weights,
                            int weights_on_device, const uint64_t r[4], const uint64_t s[4], int lane);
int zkb_prove_shard_collect(zkb_ctx* ctx, zkb_comm* comm, int lane, zkb_proof* out);
/* `count` proofs, each over all ranks, up to four in flight (zkb_prove_batch's pipelining; same results). */
int zkb_prove_shard_batch(zkb_ctx* ctx, zkb_comm* comm, const zkb_qap* qap, const zkb_crs* crs, const uint64_t* const* weights,
                          int weights_on_device, const uint64_t* r, const uint64_t* s, size_t count, zkb_proof* out);
/* One size-2^log_n transform (zkb_ntt_fr's convention) with its outer dimension sharded over the ranks: d_local holds
 * n/G canonical residues, x[rank + G*i] on entry and X[(rank*q + t) + (n/G)*k1] at index k1*q + t on return.
 * One all-to-all of n*32*(G-1)/G bytes in total and log G butterfly stages.  no_wait != 0: return once queued
 * (zkb_sync completes it). */
int zkb_ntt_shard(zkb_ctx* ctx, zkb_comm* comm, uint64_t* d_local, uint32_t log_n, int inverse, int no_wait);

/* h(x) alone: h = (u_sum * v_sum - w_sum) / t  (mod.rs:277; coefficient_poly.rs:93-157;
 * field/mod.rs:428-469).  Outputs (host, canonical, n x 4 limbs each; any may be NULL):
 * u_sum, v_sum coefficient vectors and h (n-1 meaningful coefficients, h[n-1] = 0). */
int zkb_qap_h(zkb_ctx* ctx, const zkb_qap* qap, const uint64_t* weights, uint64_t* u_sum,
              uint64_t* v_sum, uint64_t* h);

/* ---- witness generation on the device: replaces `weights()` and `evaluate()`
 * (groth16/circuit/mod.rs:529-656) and the builder's memoised `Circuit::evaluate` (circuit/builder/mod.rs:535-580).
 * The reference evaluates `(= var (* lhs rhs))` assignment by assignment; on the QAP that is, per gate k,
 *   a[out_k] = <u row of gate k, a> * <v row of gate k, a> / (w coefficient),   out_k = the single wire in gate k's w row.
 * A plan levelises the gates once per circuit (level = 1 + the deepest producer of an input; free wires and the
 * unity wire are level 0); zkb_witness_generate then runs one launch per wide level (one thread per gate) and one
 * single-block launch per run of narrow levels.  Gates with an empty w row assign nothing and are skipped (padding
 * gates of a re-indexed QAP).  free_wires: the wires whose values the caller supplies -- the `(in ...)` variables of
 * the program (circuit/mod.rs:543-567), as indices into the weight vector (1 .. m-1; wire 0 is the constant 1).
 * Errors (ZKB_ERR_ARG, message as in the reference): a wire assigned twice ("Attempted to assign to an already assigned
 * variable", :601-606), a gate reading a wire nothing produces -- with ZKB_WITNESS_PROGRAM_ORDER also one only a LATER
 * gate produces, which is what the sequential walk rejects -- ("Under constrained expression", :608-616), a wire
 * nothing assigns ("Every variable should have an assignment", :630), "Wrong number of values supplied" (:553-558).
 * Without the flag any topological order is accepted (the builder's demand-driven evaluate).  A gate with several
 * wires in its w row is ZKB_ERR_UNSUPPORTED.  The plan copies what it needs (a level-ordered CSR of the gates' u / v
 * entries): it stays valid after the QAP is freed.  Device buffers passed in (values, weights_out) must be 16-byte
 * aligned (zkb_dev_alloc's are). */
typedef struct zkb_witness_plan zkb_witness_plan;
#define ZKB_WITNESS_PROGRAM_ORDER 1
int zkb_witness_plan_create(zkb_ctx* ctx, const zkb_qap* qap, const uint32_t* free_wires, size_t n_free, int flags,
                            zkb_witness_plan** out);
/* gates that assign a wire, levels (circuit depth), widest level, kernel launches per zkb_witness_generate; any may be NULL */
int zkb_witness_plan_info(const zkb_witness_plan* plan, uint64_t* n_gates, uint64_t* n_levels, uint64_t* max_width,
                          uint64_t* n_launches);
/* values: n_values x 4 limbs canonical (host or device), in free_wires order.  weights_out: m x 4 limbs canonical --
 * `[1] ++ assignments in wire order` (circuit/mod.rs:634-636) -- in host memory, or (out_on_device) in device memory
 * where zkb_prove_dev / zkb_prove_batch(weights_on_device) read it without a host round trip. */
int zkb_witness_generate(zkb_ctx* ctx, const zkb_witness_plan* plan, const uint64_t* values, size_t n_values,
                         int values_on_device, uint64_t* weights_out, int out_on_device);
void zkb_witness_plan_free(zkb_ctx* ctx, zkb_witness_plan* plan);
/* Host only (no device, no context; errors via zkb_last_error(NULL)): the level the planner assigns to every gate of a
 * QAP in its upload form -- gate_level[k] in 1 .. *n_levels, 0 for gates that assign nothing -- with the same checks and
 * error messages as zkb_witness_plan_create. */
int zkb_witness_levels(const zkb_qap_host* qap, const uint32_t* free_wires, size_t n_free, int flags, uint32_t* gate_level,
                       uint64_t* n_levels);

/* ---- wire format: flat little-endian layout of QAP, CRS and Proof (host-only; no device is touched) -------------
 * The reference cannot serialise anything: QAP, SigmaG1, SigmaG2, Proof have private fields and no accessors
 * (groth16/mod.rs:60-128).  A prover service needs to (setup once, ship CRS + QAP to the GPU box, get proofs back), so:
 * 64-byte header (magic "ZKB200W", version, kind 1 QAP / 2 CRS / 3 Proof, total length, four size fields, FNV-1a-64 of
 * the payload) followed by the arrays of the zkb_*_host structs in declaration order, each padded to 8 bytes (layout
 * in csrc/wire.cu).  Writers fill a caller buffer of zkb_wire_size_* bytes.  Readers check magic, version, kind, length
 * and checksum and return VIEWS: the pointers of *out point into `buf`, which must be 8-byte aligned and outlive them
 * (pass the struct straight to zkb_qap_upload / zkb_crs_upload*).  ctx-less: errors via zkb_last_error(NULL). */
#define ZKB_WIRE_PROOF_BYTES 320
int zkb_wire_kind(const uint8_t* buf, uint64_t len, int* kind, uint64_t* total_bytes);
int zkb_wire_size_qap(const zkb_qap_host* qap, uint64_t* bytes);
int zkb_wire_write_qap(const zkb_qap_host* qap, uint8_t* buf, uint64_t cap);
int zkb_wire_read_qap(const uint8_t* buf, uint64_t len, zkb_qap_host* out);
int zkb_wire_size_crs(const zkb_crs_host* crs, uint64_t* bytes);
int zkb_wire_write_crs(const zkb_crs_host* crs, uint8_t* buf, uint64_t cap);
int zkb_wire_read_crs(const uint8_t* buf, uint64_t len, zkb_crs_host* out);
int zkb_wire_write_proof(const zkb_proof* proof, uint8_t* buf, uint64_t cap);
int zkb_wire_read_proof(const uint8_t* buf, uint64_t len, zkb_proof* out);

/* ---- groth16::verify (groth16/mod.rs:299-320) ----------------------------------------------- */
/* *ok = 1 iff  e(alpha1, beta2) * e(sum_term, gamma2) * e(proof.c, delta2) == e(proof.a, proof.b)
 * with sum_term = sum over zip(sigma_g1.sum_gamma, [1] ++ inputs) of a * point (mod.rs:312-316: the
 * shorter side ends the sum).  inputs: n_inputs x 4 limbs (host, canonical).  The CRS is borrowed, not
 * consumed (the reference takes it by value, mod.rs:300); a sharded CRS works on every rank (each
 * rank keeps the fixed points and sum_gamma).  Proof points that are not on their curve (impossible
 * to construct through the reference's types) give *ok = 0.  The pairing is the optimal-ate pairing
 * on BN254 (crate `bn`'s `pairing`, fr.rs:120-122); GT "+" is the Fq12 product (fr.rs:225-231). */
int zkb_verify(zkb_ctx* ctx, const zkb_crs* crs, const uint64_t* inputs, size_t n_inputs, const zkb_proof* proof, int* ok);
/* `count` proofs against the same CRS, one verdict each (inputs: count x n_inputs x 4 limbs): the
 * independent Miller loops and final exponentiations run side by side on the device. */
int zkb_verify_batch(zkb_ctx* ctx, const zkb_crs* crs, const uint64_t* inputs, size_t n_inputs,
                     const zkb_proof* proofs, size_t count, int* ok);
/* prod_i e(g1s[i], g2s[i]) as a GT element (fr.rs:120-122, 225-231): 12 Fq residues = the Fq2
 * coefficients (c0, c1) of w^0 .. w^5 in Fq12 = Fq2[w]/(w^6 - (9 + u)); 48 limbs, host.  g1s: n x 8,
 * g2s: n x 16 limbs (host); the identity in either slot contributes 1; n = 0 gives 1. */
int zkb_pairing(zkb_ctx* ctx, const uint64_t* g1s, const uint64_t* g2s, size_t n, uint64_t* gt);

/* ---- the two kernels standalone ------------------------------------------------------------- */
/* In-place size-2^log_n transform of a DEVICE vector of canonical Fr residues with the reference's
 * convention (field/mod.rs:508-537): out[i] = sum_j in[j] * root^(i*j), natural order in and out,
 * root = omega_{2^log_n} (forward) or its inverse with the 1/n scaling (inverse != 0).
 * coset_shift (host, 4 limbs, may be NULL): forward evaluates on shift*omega^i; inverse undoes it. */
int zkb_ntt_fr(zkb_ctx* ctx, uint64_t* d_data, uint32_t log_n, int inverse, const uint64_t* coset_shift);
/* Multi-GPU transform with the outer dimension sharded over G = 2^log_g ranks: rank g holds
 * the decimated subsequence x[g], x[g+G], ... and transforms it with zkb_ntt_fr (size n/G, same `inverse`);
 * the G partial transforms are all-gathered (NCCL) into d_parts (G x n/G x 4 limbs, canonical, device, rank
 * order) and every rank evaluates its slice of `count` consecutive outputs from k0:
 *   X[k] = sum_g root^(g k) * Y_g[k mod n/G]     (inverse: the inverse root and the remaining factor 1/G).
 * d_out: count x 4 limbs, canonical, device.  The concatenated slices equal zkb_ntt_fr on the whole vector. */
int zkb_ntt_combine(zkb_ctx* ctx, const uint64_t* d_parts, uint32_t log_n, uint32_t log_g, int inverse, uint64_t k0,
                    uint64_t count, uint64_t* d_out);
/* Kernel-only variant for measurement: data already in Montgomery form, natural order in,
 * bit-reversed order out (forward DIF) -- the form the prove pipeline uses internally. */
int zkb_ntt_fr_raw(zkb_ctx* ctx, uint64_t* d_data_mont, uint32_t log_n, int inverse);
/* Element-wise conversion of a device vector between canonical and Montgomery form. */
int zkb_fr_to_mont(zkb_ctx* ctx, uint64_t* d_data, size_t n, int to_mont);

/* Resident base-point vectors.  group: 1 = G1 (8 limbs / point), 2 = G2 (16 limbs / point).  The
 * points are expanded on the device into the fixed-base window table the MSM reads
 * (W = 254/c + 1 rows of 2^(c*j) multiples; c chosen from n). */
int zkb_bases_upload(zkb_ctx* ctx, int group, const uint64_t* h_points, size_t n, zkb_bases** out);
/* P_i = k_i * base (base = 69*G1::one() or 96*G2::one(), i.e. encrypt_g1 / encrypt_g2 of k_i,
 * fr.rs:106-113), computed on the device from n host scalars. */
int zkb_bases_generate(zkb_ctx* ctx, int group, const uint64_t* h_scalars, size_t n, zkb_bases** out);
int zkb_bases_download(zkb_ctx* ctx, const zkb_bases* b, uint64_t* h_points);
void zkb_bases_free(zkb_ctx* ctx, zkb_bases* b);
/* sum_i scalars[i] * bases[i]  (the `.zip().map(exp_encrypted_g*).sum()` pattern, groth16/mod.rs:
 * 255-272, 279-290; fr.rs:114-119, 191-223).  scalars: n x 4 limbs canonical, host or device.
 * n may be smaller than the base vector (zip truncation).  out: 8 (G1) or 16 (G2) limbs, host.
 * window_bits = 0 uses the table as built; another value rebuilds the table of `bases` for that
 * window size first (slow; meant for tests and tuning). */
int zkb_msm(zkb_ctx* ctx, zkb_bases* bases, const uint64_t* scalars, int scalars_on_device,
            size_t n, int window_bits, uint64_t* out);
/* Window sharding (BASELINE.json config 4): the partial sum over the table rows
 * (windows) j = win_rank (mod win_world) of the same MSM.  Every rank holds the full base vector and
 * all scalars, sorts and accumulates only its windows' records (1/win_world of the point additions),
 * and the win_world partial points -- all-gathered over NCCL -- sum to the zkb_msm result
 * (zkb_points_sum; compared in affine form, so bit-exact against one GPU). */
int zkb_msm_windows(zkb_ctx* ctx, zkb_bases* bases, const uint64_t* scalars, int scalars_on_device,
                    size_t n, int win_rank, int win_world, uint64_t* out);
/* Sum of n affine points (fold of per-GPU partial results; `Sum for G1Local`, fr.rs:191-198). */
int zkb_points_sum(zkb_ctx* ctx, int group, const uint64_t* h_points, size_t n, uint64_t* out);

/* Element-wise field operations on the device (validation hook for the field arithmetic the kernels
 * are built from: crate `bn`'s Fr / Fq / Fq2 ops reached through fr.rs:18-56).  field: 0 Fr, 1 Fq
 * (4 limbs per element), 2 Fq2 (8 limbs: c0, c1).  op for Fr/Fq: 0 a*b, 1 a^2, 2 a*b - c*d, 3 a+b,
 * 4 a-b, 5 1/a (0 -> 0), 6 / 7 a*b as wide product + stand-alone reduction (7: Karatsuba wide product); for Fq2: 0 a*b, 1 a^2, 2 1/a, 3 a*b - c*d (the two-reduction form the G2 mixed addition
 * uses).  All arrays host, canonical, n elements. */
int zkb_field_op(zkb_ctx* ctx, int field, int op, const uint64_t* a, const uint64_t* b, const uint64_t* c,
                 const uint64_t* d, uint64_t* out, size_t n);

/* Peak-rate micro-benchmark: every thread runs `iters` dependent-chain pairs of Fq Montgomery
 * multiplications (ILP 4); returns measured modmul/s in *rate (denominator of the point-add
 * roofline) and the launch duration in *ms. */
int zkb_bench_modmul(zkb_ctx* ctx, int field /*0 Fr, 1 Fq; 2 / 3: Fq as wide product + reduction, 3 with Karatsuba*/, int iters,
                     double* rate, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */
