/* Oracle F -- "best-effort CPU": the prove() path with FAST algorithms on all host threads.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  SURVEY.md 8(d) asks, beside the reference's own
 * (naive, single-threaded) algorithm that Oracle B restates and times, for a multi-threaded CPU build
 * with an NTT and Pippenger "for context": this file.  It is NOT the reference's algorithm and is not
 * the reference arm of bench.py; it answers "what would a reasonable CPU prover do on this box" so the
 * GPU number is not only compared with an O(n^2) loop.  Never linked into or called from the product.
 *
 * Same results as groth16::prove (/root/reference/src/groth16/mod.rs:213-296) on the roots-of-unity
 * domain: A, B, C are compared bit for bit with Oracle B / Oracle A in tests/test_oracle_fast.py.
 *   weighted sums (mod.rs:233-253)   ->  caller passes A_k = <a, u(w^k)>, B_k = <a, v(w^k)>; inverse NTT
 *   Mul + Div (mod.rs:277)           ->  size-2n NTT product; t = x^n - 1, so the quotient is the high
 *                                         half of u_sum * v_sum (w_sum only touches the discarded remainder)
 *   5 MSMs (mod.rs:255-290)          ->  Pippenger: windows x point slices as work units over pthreads,
 *                                         Jacobian buckets, mixed additions (madd-2007-bl)
 *   assembly (mod.rs:274-293)        ->  as Oracle B
 * Arithmetic: the primitives of oracle_b.c (included below), i.e. independent of the CUDA constants.
 */
#include "oracle_b.c"
#include <pthread.h>

/* ---- parallel-for over [0, n) in T contiguous slices ---- */
typedef void (*pf_fn)(void* arg, size_t lo, size_t hi, int tid);
typedef struct { pf_fn fn; void* arg; size_t lo, hi; int tid; } pf_job;
static void* pf_tramp(void* p) { pf_job* j = p; j->fn(j->arg, j->lo, j->hi, j->tid); return NULL; }
static void par_for(int T, size_t n, pf_fn fn, void* arg) {
  if (T < 1) T = 1;
  if ((size_t)T > n) T = n ? (int)n : 1;
  if (T == 1) { fn(arg, 0, n, 0); return; }
  pthread_t th[256]; pf_job jb[256];
  if (T > 256) T = 256;
  for (int t = 0; t < T; t++) {
    jb[t].fn = fn; jb[t].arg = arg; jb[t].tid = t;
    jb[t].lo = n * (size_t)t / T; jb[t].hi = n * (size_t)(t + 1) / T;
    pthread_create(&th[t], NULL, pf_tramp, &jb[t]);
  }
  for (int t = 0; t < T; t++) pthread_join(th[t], NULL);
}

/* ---- NTT over Fr (Montgomery form, natural order in and out) ---- */
static fe fr_omega(unsigned log_n, int inverse) {
  /* 5 is a non-residue: w_{2^28} = 5^((r-1)/2^28) */
  u64 e[4], one[4] = {1, 0, 0, 0};
  sub4(e, FR.p, one);
  for (int s = 0; s < 28; s++) { /* e >>= 1 */
    for (int i = 0; i < 4; i++) e[i] = (e[i] >> 1) | (i < 3 ? e[i + 1] << 63 : 0);
  }
  u64 five[4] = {5, 0, 0, 0};
  fe w = fe_pow(&FR, fe_from_canon(&FR, five), e);
  for (unsigned i = log_n; i < 28; i++) w = fe_sqr(&FR, w);
  return inverse ? fe_inv(&FR, w) : w;
}
typedef struct { fe* d; const fe* tw; unsigned log_n, log_blk, s0; } ntt_ctx;
static void ntt_blocks(void* a, size_t lo, size_t hi, int tid) { /* stages 0..log_blk-1 inside each block */
  (void)tid; ntt_ctx* c = a; size_t n = (size_t)1 << c->log_n;
  for (size_t b = lo; b < hi; b++) {
    fe* d = c->d + (b << c->log_blk);
    for (unsigned s = 0; s < c->log_blk; s++) {
      size_t half = (size_t)1 << s, step = n >> (s + 1);
      for (size_t g = 0; g < ((size_t)1 << c->log_blk); g += 2 * half)
        for (size_t j = 0; j < half; j++) {
          fe x = d[g + j], y = fe_mul(&FR, d[g + j + half], c->tw[j * step]);
          d[g + j] = fe_add(&FR, x, y); d[g + j + half] = fe_sub(&FR, x, y);
        }
    }
  }
}
static void ntt_stage(void* a, size_t lo, size_t hi, int tid) { /* butterflies lo..hi of stage s0 */
  (void)tid; ntt_ctx* c = a; size_t n = (size_t)1 << c->log_n;
  size_t half = (size_t)1 << c->s0, step = n >> (c->s0 + 1);
  for (size_t i = lo; i < hi; i++) {
    size_t j = i & (half - 1), p = ((i >> c->s0) << (c->s0 + 1)) + j;
    fe x = c->d[p], y = fe_mul(&FR, c->d[p + half], c->tw[j * step]);
    c->d[p] = fe_add(&FR, x, y); c->d[p + half] = fe_sub(&FR, x, y);
  }
}
typedef struct { fe* d; fe s; } scale_ctx;
static void scale_range(void* a, size_t lo, size_t hi, int tid) {
  (void)tid; scale_ctx* c = a;
  for (size_t i = lo; i < hi; i++) c->d[i] = fe_mul(&FR, c->d[i], c->s);
}
typedef struct { fe* tw; fe w; size_t chunk; } tw_ctx;
static void tw_fill(void* a, size_t lo, size_t hi, int tid) {
  (void)tid; tw_ctx* c = a;
  for (size_t b = lo; b < hi; b++) {
    size_t i0 = b * c->chunk; u64 e[4] = {i0, 0, 0, 0};
    fe v = fe_pow(&FR, c->w, e);
    for (size_t i = 0; i < c->chunk; i++) { c->tw[i0 + i] = v; v = fe_mul(&FR, v, c->w); }
  }
}
static void ntt_mont(fe* d, unsigned log_n, int inverse, int T) {
  if (log_n == 0) return;
  size_t n = (size_t)1 << log_n, half = n >> 1;
  fe* tw = malloc(half * sizeof(fe));
  tw_ctx tc = {tw, fr_omega(log_n, inverse), half >= 1024 ? 1024 : half};
  par_for(T, half / tc.chunk, tw_fill, &tc);
  for (size_t i = 0; i < n; i++) { /* bit reversal */
    size_t j = 0; for (unsigned b = 0; b < log_n; b++) j |= ((i >> b) & 1) << (log_n - 1 - b);
    if (j > i) { fe t = d[i]; d[i] = d[j]; d[j] = t; }
  }
  ntt_ctx c = {d, tw, log_n, log_n < 12 ? log_n : 12, 0};
  par_for(T, n >> c.log_blk, ntt_blocks, &c);
  for (unsigned s = c.log_blk; s < log_n; s++) { c.s0 = s; par_for(T, half, ntt_stage, &c); }
  if (inverse) {
    u64 nn[4] = {n, 0, 0, 0};
    scale_ctx sc = {d, fe_inv(&FR, fe_from_canon(&FR, nn))};
    par_for(T, n, scale_range, &sc);
  }
  free(tw);
}
/* dft / idft convention of field/mod.rs:508-537: X[i] = sum_j x[j] root^(i j); idft scales by 1/n */
void of_ntt(u64* data, unsigned log_n, int inverse, int threads) {
  ob_init();
  size_t n = (size_t)1 << log_n;
  fe* d = malloc(n * sizeof(fe));
  for (size_t i = 0; i < n; i++) d[i] = fe_from_canon(&FR, data + 4 * i);
  ntt_mont(d, log_n, inverse, threads);
  for (size_t i = 0; i < n; i++) fe_to_canon(&FR, d[i], data + 4 * i);
  free(d);
}

/* ---- Pippenger ---- */
typedef struct { fe x, y; } aff1;   /* Montgomery; identity = (0, 0) */
typedef struct { fe2 x, y; } aff2;
#define DEFINE_MSM(G, A, T, ADD, SUB, MUL, SQR, ISZ, EQ, ONE)                                        \
  static G G##_madd(G p, A q) { /* madd-2007-bl; q affine, identity = (0,0) */                       \
    if (ISZ(q.x) && ISZ(q.y)) return p;                                                              \
    if (G##_is_zero(p)) { G r; r.x = q.x; r.y = q.y; r.z = ONE; return r; }                          \
    T z1z1 = SQR(p.z), u2 = MUL(q.x, z1z1), s2 = MUL(MUL(q.y, p.z), z1z1);                           \
    if (EQ(u2, p.x)) {                                                                               \
      if (EQ(s2, p.y)) { G r; r.x = q.x; r.y = q.y; r.z = ONE; return G##_dbl(r); }                  \
      return G##_zero();                                                                             \
    }                                                                                                \
    T h = SUB(u2, p.x), hh = SQR(h), i = ADD(hh, hh); i = ADD(i, i);                                 \
    T j = MUL(h, i), rr = SUB(s2, p.y); rr = ADD(rr, rr);                                            \
    T v = MUL(p.x, i);                                                                               \
    G r; r.x = SUB(SUB(SQR(rr), j), ADD(v, v));                                                      \
    T yj = MUL(p.y, j); yj = ADD(yj, yj);                                                            \
    r.y = SUB(MUL(rr, SUB(v, r.x)), yj);                                                             \
    r.z = SUB(SUB(SQR(ADD(p.z, h)), z1z1), hh);                                                      \
    return r;                                                                                        \
  }                                                                                                  \
  typedef struct { const A* pts; const u64* sc; size_t n; unsigned c, W, S; G* part; } G##_pip;      \
  static void G##_unit(void* a, size_t lo, size_t hi, int tid) {                                     \
    (void)tid; G##_pip* P = a;                                                                       \
    size_t nb = ((size_t)1 << P->c) - 1;                                                             \
    G* bk = malloc((nb + 1) * sizeof(G));                                                            \
    for (size_t u = lo; u < hi; u++) {                                                               \
      unsigned w = (unsigned)(u / P->S), s = (unsigned)(u % P->S);                                   \
      size_t i0 = P->n * s / P->S, i1 = P->n * (s + 1) / P->S;                                       \
      for (size_t b = 0; b <= nb; b++) bk[b] = G##_zero();                                           \
      unsigned bit = w * P->c;                                                                       \
      for (size_t i = i0; i < i1; i++) {                                                             \
        const u64* k = P->sc + 4 * i;                                                                \
        u64 d = k[bit >> 6] >> (bit & 63);                                                           \
        if ((bit & 63) + P->c > 64 && (bit >> 6) < 3) d |= k[(bit >> 6) + 1] << (64 - (bit & 63));    \
        d &= nb;                                                                                     \
        if (d) bk[d] = G##_madd(bk[d], P->pts[i]);                                                   \
      }                                                                                              \
      G run = G##_zero(), acc = G##_zero();                                                          \
      for (size_t b = nb; b >= 1; b--) { run = G##_add(run, bk[b]); acc = G##_add(acc, run); }       \
      P->part[u] = acc;                                                                              \
    }                                                                                                \
    free(bk);                                                                                        \
  }                                                                                                  \
  static G G##_pippenger(const u64* sc /* canonical limbs */, const A* pts, size_t n, int T) {       \
    if (n == 0) return G##_zero();                                                                   \
    if (T < 1) T = 1;                                                                                \
    unsigned bc = 1, bS = 1; double best = 1e300;                                                    \
    for (unsigned c = 1; c <= 20; c++)                                                               \
      for (unsigned S = 1; S <= (unsigned)T; S++) {                                                  \
        unsigned W = (254 + c - 1) / c;                                                              \
        double rounds = (double)((W * S + T - 1) / T);                                               \
        double cost = rounds * ((double)n / S * 11.0 + (double)((size_t)1 << c) * 32.0);             \
        if (cost < best) { best = cost; bc = c; bS = S; }                                            \
      }                                                                                              \
    G##_pip P; P.pts = pts; P.sc = sc; P.n = n; P.c = bc; P.W = (254 + bc - 1) / bc; P.S = bS;       \
    size_t units = (size_t)P.W * P.S;                                                                \
    P.part = malloc(units * sizeof(G));                                                              \
    par_for(T, units, G##_unit, &P);                                                                 \
    G acc = G##_zero();                                                                              \
    for (int w = (int)P.W - 1; w >= 0; w--) {                                                        \
      for (unsigned k = 0; k < bc; k++) acc = G##_dbl(acc);                                          \
      for (unsigned s = 0; s < P.S; s++) acc = G##_add(acc, P.part[(size_t)w * P.S + s]);            \
    }                                                                                                \
    free(P.part);                                                                                    \
    return acc;                                                                                      \
  }

DEFINE_MSM(g1, aff1, fe, q_add, q_sub, q_mul, q_sqr, fe_is_zero, fe_eq, Q_ONE)
DEFINE_MSM(g2, aff2, fe2, f2_add, f2_sub, f2_mul, f2_sqr, f2_is_zero, f2_eq, F2_ONE)

static aff1* load_g1(const u64* p, size_t n) {
  aff1* a = malloc((n + 1) * sizeof(aff1));
  for (size_t i = 0; i < n; i++) { a[i].x = fe_from_canon(&FQ, p + 8 * i); a[i].y = fe_from_canon(&FQ, p + 8 * i + 4); }
  return a;
}
static aff2* load_g2(const u64* p, size_t n) {
  aff2* a = malloc((n + 1) * sizeof(aff2));
  for (size_t i = 0; i < n; i++) {
    a[i].x.c0 = fe_from_canon(&FQ, p + 16 * i); a[i].x.c1 = fe_from_canon(&FQ, p + 16 * i + 4);
    a[i].y.c0 = fe_from_canon(&FQ, p + 16 * i + 8); a[i].y.c1 = fe_from_canon(&FQ, p + 16 * i + 12);
  }
  return a;
}
void of_msm_g1(const u64* scalars, const u64* pts, size_t n, int threads, u64* out) {
  ob_init();
  aff1* a = load_g1(pts, n);
  g1_store(g1_pippenger(scalars, a, n, threads), out); free(a);
}
void of_msm_g2(const u64* scalars, const u64* pts, size_t n, int threads, u64* out) {
  ob_init();
  aff2* a = load_g2(pts, n);
  g2_store(g2_pippenger(scalars, a, n, threads), out); free(a);
}

/* ---- prove on the roots-of-unity domain ---- */
typedef struct {
  size_t n;        /* gates (power of two) = xi1 / xi2 length; xi_t has n-1 entries */
  size_t n_input;  /* qap.input */
  size_t n_sd;     /* sum_delta length */
  const u64 *alpha1, *beta1, *delta1, *xi1, *xi_t, *sum_delta, *beta2, *delta2, *xi2;
} of_prove_in;

typedef struct {
  size_t n, n_sd, n_input;
  g1 alpha1, beta1, delta1; g2 beta2, delta2;
  const aff1 *xi1, *xi_t, *sd; const aff2* xi2;
} crs_int;
typedef struct { const fe* a; u64* out; } canon_ctx;
static void canon_range(void* a, size_t lo, size_t hi, int tid) {
  (void)tid; canon_ctx* c = a;
  for (size_t i = lo; i < hi; i++) fe_to_canon(&FR, c->a[i], c->out + 4 * i);
}
typedef struct { fe* a; const fe* b; } mulv_ctx;
static void mulv_range(void* a, size_t lo, size_t hi, int tid) {
  (void)tid; mulv_ctx* c = a;
  for (size_t i = lo; i < hi; i++) c->a[i] = fe_mul(&FR, c->a[i], c->b[i]);
}
/* A, B: n evaluations (Montgomery, overwritten with u_sum, v_sum); wts: nw weights (Montgomery);
 * h_out: n-1 coefficients (Montgomery) or NULL; tm[3]: seconds in polynomial stage, G1 MSMs, G2 MSM */
static void prove_core(const crs_int* C, fe* A, fe* B, const fe* wts, size_t nw, fe r, fe s, int T, g1* pa, g2* pb, g1* pc,
                       fe* h_out, double* tm) {
  size_t n = C->n; unsigned log_n = 0; while (((size_t)1 << log_n) < n) log_n++;
  double t0 = now();
  ntt_mont(A, log_n, 1, T); ntt_mont(B, log_n, 1, T);
  fe* pu = calloc(2 * n, sizeof(fe)); fe* pv = calloc(2 * n, sizeof(fe));
  memcpy(pu, A, n * sizeof(fe)); memcpy(pv, B, n * sizeof(fe));
  ntt_mont(pu, log_n + 1, 0, T); ntt_mont(pv, log_n + 1, 0, T);
  mulv_ctx mc = {pu, pv}; par_for(T, 2 * n, mulv_range, &mc);
  ntt_mont(pu, log_n + 1, 1, T);
  const fe* h = pu + n; /* n-1 coefficients (index 2n-1 is zero) */
  if (h_out) memcpy(h_out, h, (n - 1) * sizeof(fe));
  size_t skip = C->n_input + 1;
  size_t kw = nw > skip ? nw - skip : 0; if (kw > C->n_sd) kw = C->n_sd;
  size_t big = n > kw ? n : kw;
  u64* sc = malloc((big + 1) * 32);
  double t1 = now();
  canon_ctx cc = {A, sc}; par_for(T, n, canon_range, &cc);
  g1 a_g1 = g1_pippenger(sc, C->xi1, n, T);
  cc.a = B; par_for(T, n, canon_range, &cc);
  g1 b_g1 = g1_pippenger(sc, C->xi1, n, T);
  double t2 = now();
  g2 b_g2 = g2_pippenger(sc, C->xi2, n, T);
  double t3 = now();
  cc.a = h; par_for(T, n - 1, canon_range, &cc);
  g1 c = g1_pippenger(sc, C->xi_t, n - 1, T);
  cc.a = wts + (nw > skip ? skip : nw); par_for(T, kw, canon_range, &cc);
  c = g1_add(c, g1_pippenger(sc, C->sd, kw, T));
  u64 rk[4], sk[4], rsk[4];
  fe_to_canon(&FR, r, rk); fe_to_canon(&FR, s, sk); fe_to_canon(&FR, fe_mul(&FR, r, s), rsk);
  g1 a = g1_add(g1_add(a_g1, C->alpha1), g1_mul(C->delta1, rk));
  g2 b = g2_add(g2_add(b_g2, C->beta2), g2_mul(C->delta2, sk));
  c = g1_add(c, g1_mul(a, sk));
  c = g1_add(c, g1_mul(g1_add(g1_add(C->beta1, b_g1), g1_mul(C->delta1, sk)), rk));
  c = g1_add(c, g1_neg(g1_mul(C->delta1, rsk)));
  double t4 = now();
  *pa = a; *pb = b; *pc = c;
  if (tm) { tm[0] = t1 - t0; tm[1] = (t2 - t1) + (t4 - t3); tm[2] = t3 - t2; }
  free(pu); free(pv); free(sc);
}

/* A_evals, B_evals: n x 4 canonical limbs (A_k = sum_i a_i u_i(w^k), same for v); proof: 32 limbs (a | b | c);
 * h_out: (n-1) x 4 limbs or NULL.  Returns 0, or -1 if n is not a power of two >= 2. */
int of_prove(const of_prove_in* in, const u64* A_evals, const u64* B_evals, const u64* weights, size_t nw, const u64* r_,
             const u64* s_, u64* proof, u64* h_out, int threads) {
  ob_init();
  size_t n = in->n;
  if (n < 2 || (n & (n - 1))) return -1;
  fe* A = malloc(n * sizeof(fe)); fe* B = malloc(n * sizeof(fe)); fe* wts = malloc((nw + 1) * sizeof(fe));
  fe* h = malloc(n * sizeof(fe));
  for (size_t i = 0; i < n; i++) { A[i] = fe_from_canon(&FR, A_evals + 4 * i); B[i] = fe_from_canon(&FR, B_evals + 4 * i); }
  for (size_t i = 0; i < nw; i++) wts[i] = fe_from_canon(&FR, weights + 4 * i);
  aff1 *xi1 = load_g1(in->xi1, n), *xit = load_g1(in->xi_t, n - 1), *sd = load_g1(in->sum_delta, in->n_sd);
  aff2* xi2 = load_g2(in->xi2, n);
  crs_int C = {n, in->n_sd, in->n_input, g1_from_affine(in->alpha1), g1_from_affine(in->beta1), g1_from_affine(in->delta1),
               g2_from_affine(in->beta2), g2_from_affine(in->delta2), xi1, xit, sd, xi2};
  g1 a, c; g2 b;
  prove_core(&C, A, B, wts, nw, fe_from_canon(&FR, r_), fe_from_canon(&FR, s_), threads, &a, &b, &c, h, NULL);
  g1_store(a, proof); g2_store(b, proof + 8); g1_store(c, proof + 24);
  if (h_out) for (size_t i = 0; i + 1 < n; i++) fe_to_canon(&FR, h[i], h_out + 4 * i);
  free(A); free(B); free(wts); free(h); free(xi1); free(xit); free(sd); free(xi2);
  return 0;
}

/* ---- array-level entry points for the full-size parity tests (tests/test_gpu_parity_large.py) ----
 * The weighted sums of mod.rs:233-253 on the root domain: A_k = sum_i a_i u_i(w^k), B_k likewise, from the by-wire CSR rows
 * (row_ptr[m+1], gate[nnz], coeff[nnz x 4] canonical) and nw weights; wires >= nw count as zero (zip truncation). */
void of_qap_evals(size_t m, size_t n, const u64* ptr_u, const uint32_t* gate_u, const u64* coef_u, const u64* ptr_v,
                  const uint32_t* gate_v, const u64* coef_v, const u64* weights, size_t nw, u64* A_out, u64* B_out) {
  ob_init();
  const u64* ptr[2] = {ptr_u, ptr_v}; const uint32_t* gate[2] = {gate_u, gate_v}; const u64* coef[2] = {coef_u, coef_v};
  u64* out[2] = {A_out, B_out};
  fe* acc = malloc(n * sizeof(fe));
  for (int t = 0; t < 2; t++) {
    memset(acc, 0, n * sizeof(fe));
    for (size_t i = 0; i < m && i < nw; i++) {
      fe w = fe_from_canon(&FR, weights + 4 * i);
      for (u64 e = ptr[t][i]; e < ptr[t][i + 1]; e++)
        acc[gate[t][e]] = fe_add(&FR, acc[gate[t][e]], fe_mul(&FR, fe_from_canon(&FR, coef[t] + 4 * e), w));
    }
    for (size_t k = 0; k < n; k++) fe_to_canon(&FR, acc[k], out[t] + 4 * k);
  }
  free(acc);
}
/* u_sum, v_sum (n coefficients each) and h = quotient of u_sum * v_sum by x^n - 1 (n - 1 coefficients, h_out[n-1] = 0):
 * inverse NTTs + one size-2n product, the polynomial stage of prove_core.  All arrays n x 4 canonical limbs. */
int of_qap_h(const u64* A_evals, const u64* B_evals, unsigned log_n, u64* u_out, u64* v_out, u64* h_out, int threads) {
  ob_init();
  size_t n = (size_t)1 << log_n;
  if (log_n < 1) return -1;
  fe* A = malloc(n * sizeof(fe)); fe* B = malloc(n * sizeof(fe));
  for (size_t i = 0; i < n; i++) { A[i] = fe_from_canon(&FR, A_evals + 4 * i); B[i] = fe_from_canon(&FR, B_evals + 4 * i); }
  ntt_mont(A, log_n, 1, threads); ntt_mont(B, log_n, 1, threads);
  fe* pu = calloc(2 * n, sizeof(fe)); fe* pv = calloc(2 * n, sizeof(fe));
  memcpy(pu, A, n * sizeof(fe)); memcpy(pv, B, n * sizeof(fe));
  ntt_mont(pu, log_n + 1, 0, threads); ntt_mont(pv, log_n + 1, 0, threads);
  mulv_ctx mc = {pu, pv}; par_for(threads, 2 * n, mulv_range, &mc);
  ntt_mont(pu, log_n + 1, 1, threads);
  for (size_t i = 0; i < n; i++) {
    if (u_out) fe_to_canon(&FR, A[i], u_out + 4 * i);
    if (v_out) fe_to_canon(&FR, B[i], v_out + 4 * i);
    if (h_out) fe_to_canon(&FR, pu[n + i], h_out + 4 * i);  /* index 2n-1 is zero: deg(u v) <= 2n-2 */
  }
  free(A); free(B); free(pu); free(pv);
  return 0;
}
/* data[i] *= g^i (the coset shift of a forward transform) */
void of_scale_powers(u64* data, size_t n, const u64* g_) {
  ob_init();
  fe g = fe_from_canon(&FR, g_), p = FR.one;
  for (size_t i = 0; i < n; i++) {
    fe_to_canon(&FR, fe_mul(&FR, fe_from_canon(&FR, data + 4 * i), p), data + 4 * i);
    p = fe_mul(&FR, p, g);
  }
}

/* ---- timing of one full-size proof (bench.py: cpu_best_effort) ----
 * The bases are n distinct points P0 + i*D (batch-normalised), not a real CRS: the cost of an MSM
 * does not depend on which points it folds.  Scalars, evaluations and weights are uniform in Fr. */
typedef struct { aff1* out; size_t n; u64 seed; } gen1_ctx;
#define GEN_CHUNK 4096
static void gen1_chunk(gen1_ctx* c, size_t lo, size_t hi) {
  u64 st = c->seed + 977 * (u64)(lo / GEN_CHUNK + 1), k[4];
  u64 gen[8] = {1, 0, 0, 0, 2, 0, 0, 0};
  fe_to_canon(&FR, rand_fr(&st), k); g1 P = g1_mul(g1_from_affine(gen), k);
  fe_to_canon(&FR, rand_fr(&st), k); g1 D = g1_mul(g1_from_affine(gen), k);
  size_t m = hi - lo;
  g1* J = malloc(m * sizeof(g1)); fe* pre = malloc(m * sizeof(fe));
  fe run = FQ.one;
  for (size_t i = 0; i < m; i++) { J[i] = P; P = g1_add(P, D); pre[i] = run; run = q_mul(run, J[i].z); }
  fe inv = q_inv(run);
  for (size_t i = m; i-- > 0;) {
    fe zi = q_mul(inv, pre[i]); inv = q_mul(inv, J[i].z);
    fe zi2 = q_sqr(zi);
    c->out[lo + i].x = q_mul(J[i].x, zi2); c->out[lo + i].y = q_mul(J[i].y, q_mul(zi2, zi));
  }
  free(J); free(pre);
}
static void gen1_range(void* a, size_t lo, size_t hi, int tid) { /* lo, hi count chunks */
  (void)tid; gen1_ctx* c = a;
  for (size_t b = lo; b < hi; b++) { size_t e = (b + 1) * GEN_CHUNK; gen1_chunk(c, b * GEN_CHUNK, e < c->n ? e : c->n); }
}
typedef struct { aff2* out; size_t n; u64 seed; const u64* g2gen; } gen2_ctx;
static void gen2_chunk(gen2_ctx* c, size_t lo, size_t hi) {
  u64 st = c->seed + 1913 * (u64)(lo / GEN_CHUNK + 1), k[4];
  fe_to_canon(&FR, rand_fr(&st), k); g2 P = g2_mul(g2_from_affine(c->g2gen), k);
  fe_to_canon(&FR, rand_fr(&st), k); g2 D = g2_mul(g2_from_affine(c->g2gen), k);
  size_t m = hi - lo;
  g2* J = malloc(m * sizeof(g2)); fe2* pre = malloc(m * sizeof(fe2));
  fe2 run = f2_one();
  for (size_t i = 0; i < m; i++) { J[i] = P; P = g2_add(P, D); pre[i] = run; run = f2_mul(run, J[i].z); }
  fe2 inv = f2_inv(run);
  for (size_t i = m; i-- > 0;) {
    fe2 zi = f2_mul(inv, pre[i]); inv = f2_mul(inv, J[i].z);
    fe2 zi2 = f2_sqr(zi);
    c->out[lo + i].x = f2_mul(J[i].x, zi2); c->out[lo + i].y = f2_mul(J[i].y, f2_mul(zi2, zi));
  }
  free(J); free(pre);
}
static void gen2_range(void* a, size_t lo, size_t hi, int tid) {
  (void)tid; gen2_ctx* c = a;
  for (size_t b = lo; b < hi; b++) { size_t e = (b + 1) * GEN_CHUNK; gen2_chunk(c, b * GEN_CHUNK, e < c->n ? e : c->n); }
}
typedef struct { fe* d; u64 seed; size_t n; } rnd_ctx;
static void rnd_range(void* a, size_t lo, size_t hi, int tid) { /* lo, hi count chunks of GEN_CHUNK elements */
  (void)tid; rnd_ctx* c = a;
  for (size_t b = lo; b < hi; b++) {
    u64 st = c->seed ^ (0x5bd1e995ull * (u64)(b + 1));
    for (size_t i = b * GEN_CHUNK; i < (b + 1) * GEN_CHUNK && i < c->n; i++) c->d[i] = rand_fr(&st);
  }
}
/* Returns seconds for ONE proof at n = 2^log_n, m = 2n+2, input = 2 (the synthetic Horner shape);
 * tm[0..2] = polynomial stage, G1 MSMs (5n-1 terms), G2 MSM (n terms); check[0..3] = limbs of proof.a.x
 * (so the work cannot be optimised away and two runs can be compared). */
double of_time_prove(unsigned log_n, int threads, u64 seed, const u64* g2gen, double* tm, u64* check) {
  ob_init();
  size_t n = (size_t)1 << log_n, m = 2 * n + 2, nsd = m - 3;
  int T = threads;
  aff1* g1pts = malloc((2 * n + nsd) * sizeof(aff1)); aff2* g2pts = malloc(n * sizeof(aff2));
#define NCHUNK(x) (((x) + GEN_CHUNK - 1) / GEN_CHUNK)
  gen1_ctx c1 = {g1pts, 2 * n + nsd, seed}; par_for(T, NCHUNK(2 * n + nsd), gen1_range, &c1);
  gen2_ctx c2 = {g2pts, n, seed + 1, g2gen}; par_for(T, NCHUNK(n), gen2_range, &c2);
  fe* A = malloc(n * sizeof(fe)); fe* B = malloc(n * sizeof(fe)); fe* wts = malloc(m * sizeof(fe));
  rnd_ctx r1 = {A, seed + 2, n}, r2 = {B, seed + 3, n}, r3 = {wts, seed + 4, m};
  par_for(T, NCHUNK(n), rnd_range, &r1); par_for(T, NCHUNK(n), rnd_range, &r2); par_for(T, NCHUNK(m), rnd_range, &r3);
  u64 st = seed + 5;
  fe r = rand_fr(&st), s = rand_fr(&st);
  crs_int C = {n, nsd, 2, {g1pts[0].x, g1pts[0].y, FQ.one}, {g1pts[1].x, g1pts[1].y, FQ.one}, {g1pts[2].x, g1pts[2].y, FQ.one},
               {g2pts[0].x, g2pts[0].y, f2_one()}, {g2pts[1].x, g2pts[1].y, f2_one()}, g1pts, g1pts + n, g1pts + 2 * n, g2pts};
  g1 a, c; g2 b;
  double t0 = now();
  prove_core(&C, A, B, wts, m, r, s, T, &a, &b, &c, NULL, tm);
  double t = now() - t0;
  u64 pa[8]; g1_store(a, pa);
  if (check) memcpy(check, pa, 32);
  free(g1pts); free(g2pts); free(A); free(B); free(wts);
  return t;
}
