// The reference's own BN254 tests (src/groth16/fr.rs:248-416, src/groth16/mod.rs:635-690) rewritten against the C++
// host mirror include/zkb200.hpp, so that they read like the originals: build a QAP from a root representation,
// setup, prove, assert verify(...) == true under fresh randomness.  A last section prints one proof made from fixed
// secrets so that tests/test_cpp_host.py can compare it bit for bit with the oracle.
// Exit codes: 0 ok, 1 assertion failed, 3 zkb200::Error (e.g. no CUDA device: the library has no CPU fallback).
#include <cstdio>
#include "zkb200.hpp"
using namespace zkb200;

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) { printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

static void print_limbs(const char* name, const uint64_t* v, int n) {
  printf("%s", name);
  for (int i = 0; i < n; i++) printf(" %016llx", (unsigned long long)v[i]);
  printf("\n");
}

// fr.rs:248-271: one multiplication gate, t = x + 250 (root -250), weights [1, 51, 3, 17], inputs [51, 3]
static int single_mult_honest_bn(Context& ctx) {
  Fr root = -Fr(250), one = Fr::one();
  RootRepresentation rep;
  rep.u = {{}, {}, {{root, one}}, {}};
  rep.v = {{}, {}, {}, {{root, one}}};
  rep.w = {{}, {{root, one}}, {}, {}};
  rep.roots = {root};
  rep.input = 2;
  QAP qap = QAP::from(ctx, rep);
  std::vector<Fr> weights = {1, 51, 3, 17};
  for (int i = 0; i < 10; i++) {
    Sigma sigma = groth16::setup(ctx, qap);
    Proof proof = groth16::prove(ctx, qap, sigma, weights);
    CHECK(groth16::verify(ctx, sigma, {Fr(51), Fr(3)}, proof));
    CHECK(!groth16::verify(ctx, sigma, {Fr(51), Fr(4)}, proof));
  }
  printf("ok single_mult_honest_bn\n");
  return 0;
}

// mod.rs:635-690 (qap_from_roots): quadratic share a x^2 + b x + c on the roots 1, 2, 3
static int qap_from_roots(Context& ctx) {
  Fr one = Fr::one(), r1 = 1, r2 = 2, r3 = 3;
  RootRepresentation rep;
  rep.u = {{{r3, one}}, {{r1, one}, {r2, one}}, {}, {}, {}, {}, {}, {}};
  rep.v = {{}, {}, {}, {{r1, one}}, {{r2, one}}, {{r3, one}}, {{r2, one}}, {{r3, one}}};
  rep.w = {{}, {}, {{r3, one}}, {}, {}, {}, {{r1, one}}, {{r2, one}}};
  rep.roots = {r1, r2, r3};
  rep.input = 2;
  QAP qap = QAP::from(ctx, rep);
  for (int i = 0; i < 10; i++) {
    Fr x = Fr::random_elem(), a = Fr::random_elem(), b = Fr::random_elem(), c = Fr::random_elem();
    Fr share = a * x * x + b * x + c;
    std::vector<Fr> weights = {1, x, share, a, b, c, a * x, x * (a * x + b)};
    Sigma sigma = groth16::setup(ctx, qap);
    Proof proof = groth16::prove(ctx, qap, sigma, weights);
    CHECK(groth16::verify(ctx, sigma, {x, share}, proof));
    CHECK(!groth16::verify(ctx, sigma, {x, share + one}, proof));
  }
  printf("ok qap_from_roots\n");
  return 0;
}

// the n-gate Horner circuit (test_programs/deg_15.zk for n = 16; fr.rs:361-416) on the roots of unity, rows in the
// order ASTParser produces: 0 unity, 1 x, 2 y, t_k -> 2k+1, c_k -> 2k+2 (k < n), c_n -> 2n+1
static RootRepresentation horner_rep(size_t n, int log_n) {
  RootRepresentation rep;
  const size_t m = 2 * n + 2;
  rep.u.resize(m); rep.v.resize(m); rep.w.resize(m);
  Fr w = QAP::omega(log_n), g = Fr::one(), one = Fr::one();
  for (size_t k = 1; k <= n; k++) {
    rep.roots.push_back(g);
    if (k < n) { rep.u[1].push_back({g, one}); rep.w[2 * k + 1].push_back({g, one}); }
    else { rep.u[0].push_back({g, one}); rep.w[2].push_back({g, one}); }
    if (k > 1) rep.v[2 * (k - 1) + 1].push_back({g, one});
    rep.v[k < n ? 2 * k + 2 : 2 * n + 1].push_back({g, one});
    g = g * w;
  }
  rep.input = 2;
  return rep;
}
static std::vector<Fr> horner_weights(size_t n, const Fr& x, const std::vector<Fr>& c) {
  std::vector<Fr> a(2 * n + 2);
  a[0] = Fr::one(); a[1] = x;
  Fr acc = x * c[0];
  a[3] = acc; a[4] = c[0];
  for (size_t k = 2; k < n; k++) { acc = x * (acc + c[k - 1]); a[2 * k + 1] = acc; a[2 * k + 2] = c[k - 1]; }
  a[2 * n + 1] = c[n - 1];
  a[2] = acc + c[n - 1];
  return a;
}
static int bn_encrypt_deg_15_test(Context& ctx) {
  const size_t n = 16;
  QAP qap = QAP::from(ctx, horner_rep(n, 4));
  for (int i = 0; i < 10; i++) {
    std::vector<Fr> c(n);
    for (auto& v : c) v = Fr::random_elem();
    std::vector<Fr> weights = horner_weights(n, Fr::random_elem(), c);
    Sigma sigma = groth16::setup(ctx, qap);
    Proof proof = groth16::prove(ctx, qap, sigma, weights);
    CHECK(groth16::verify(ctx, sigma, {weights[1], weights[2]}, proof));
    CHECK(!groth16::verify(ctx, sigma, {weights[1], weights[2] + Fr::one()}, proof));
  }
  printf("ok bn_encrypt_deg_15_test\n");
  return 0;
}

// throughput mode: prove_many == prove_with_rs in a loop, verify_many gives one verdict per proof
static int batch_equals_single(Context& ctx) {
  const size_t n = 16, count = 5;
  QAP qap = QAP::from(ctx, horner_rep(n, 4));
  Sigma sigma = groth16::setup(ctx, qap);
  std::vector<std::vector<Fr>> weights, inputs;
  std::vector<std::pair<Fr, Fr>> rs;
  std::vector<Proof> single;
  for (size_t i = 0; i < count; i++) {
    std::vector<Fr> c(n);
    for (auto& v : c) v = Fr::random_elem();
    weights.push_back(horner_weights(n, Fr::random_elem(), c));
    rs.push_back({Fr::random_elem(), Fr::random_elem()});
    inputs.push_back({weights[i][1], weights[i][2]});
    single.push_back(groth16::prove_with_rs(ctx, qap, sigma, weights[i], rs[i].first, rs[i].second));
  }
  std::vector<Proof> many = groth16::prove_many(ctx, qap, sigma, weights, rs);
  CHECK(many.size() == count);
  for (size_t i = 0; i < count; i++) CHECK(many[i].a == single[i].a && many[i].b == single[i].b && many[i].c == single[i].c);
  inputs[3][1] = inputs[3][1] + Fr::one();  // one wrong public input
  std::vector<bool> ok = groth16::verify_many(ctx, sigma, inputs, many);
  for (size_t i = 0; i < count; i++) CHECK(ok[i] == (i != 3));
  CHECK(groth16::prove_many(ctx, qap, sigma, {}, {}).empty() && groth16::verify_many(ctx, sigma, {}, {}).empty());
  printf("ok batch_equals_single\n");
  return 0;
}

// fixed secrets -> one proof, printed for the bit-exact comparison with the oracle (tests/test_cpp_host.py)
static int parity_dump(Context& ctx) {
  const size_t n = 8;
  QAP qap = QAP::from(ctx, horner_rep(n, 3));
  std::vector<Fr> c(n);
  for (size_t k = 0; k < n; k++) c[k] = Fr::from_str("1000000007") * Fr(k + 3) + Fr(k);
  std::vector<Fr> weights = horner_weights(n, Fr::from_str("123456789012345678901234567890"), c);
  Sigma sigma = groth16::setup_with(ctx, qap, {Fr(3), Fr(5), Fr(7), Fr(11), Fr(13)});
  Proof p = groth16::prove_with_rs(ctx, qap, sigma, weights, Fr(17), Fr(19));
  CHECK(groth16::verify(ctx, sigma, {weights[1], weights[2]}, p));
  print_limbs("proof.a", p.a.v, 8);
  print_limbs("proof.b", p.b.v, 16);
  print_limbs("proof.c", p.c.v, 8);
  // host Fr arithmetic used above, pinned: (r - 1) * (r - 1) == 1, 1/7 * 7 == 1, from_str round trip
  Fr m1 = -Fr::one();
  CHECK(m1 * m1 == Fr::one() && Fr(7).mul_inv() * Fr(7) == Fr::one() && (Fr(5) - Fr(7)) + Fr(2) == Fr::zero());
  CHECK(Fr::from_str("21888242871839275222246405745257275088548364400416034343698204186575808495616") == m1);
  printf("ok parity_dump\n");
  return 0;
}

// circuit/mod.rs:746-769 (weights_test) on the rows its program parses to (roots 1, 2; wires 1 b x temp a c), then
// weights -> prove -> verify as lib.rs:156-190 does for the same program
static int weights_test(Context& ctx) {
  Fr one = Fr::one(), r1 = 1, r2 = 2;
  RootRepresentation rep;
  rep.u = {{{r2, one}}, {}, {}, {}, {{r1, one}}, {}};                  // gate 1: a * b ; gate 2: 1 * (4 temp + c + 6)
  rep.v = {{{r2, Fr(6)}}, {{r1, one}}, {}, {{r2, Fr(4)}}, {}, {{r2, one}}};
  rep.w = {{}, {}, {{r2, one}}, {{r1, one}}, {}, {}};
  rep.roots = {r1, r2};
  rep.input = 2;
  QAP qap = QAP::from(ctx, rep);
  groth16::WitnessPlan plan(ctx, qap, {4, 1, 5});  // (in a b c)
  std::vector<Fr> w = plan.weights({Fr(3), Fr(2), Fr(4)});
  std::vector<Fr> expected = {1, 2, 34, 6, 3, 4};
  CHECK(w == expected && plan.levels() == 2);
  Sigma sigma = groth16::setup(ctx, qap);
  Proof proof = groth16::prove(ctx, qap, sigma, w);
  CHECK(groth16::verify(ctx, sigma, {w[1], w[2]}, proof));
  bool threw = false;
  try { plan.weights({Fr(3), Fr(2)}); } catch (const Error& e) { threw = e.code == ZKB_ERR_ARG; }  // "Wrong number of values supplied"
  CHECK(threw);
  printf("ok weights_test\n");
  return 0;
}

int main() {
  try {
    Context ctx(0);
    if (int rc = single_mult_honest_bn(ctx)) return rc;
    if (int rc = qap_from_roots(ctx)) return rc;
    if (int rc = bn_encrypt_deg_15_test(ctx)) return rc;
    if (int rc = batch_equals_single(ctx)) return rc;
    if (int rc = weights_test(ctx)) return rc;
    if (int rc = parity_dump(ctx)) return rc;
  } catch (const Error& e) {
    printf("zkb200::Error %d: %s\n", e.code, e.what());
    return 3;
  }
  printf("all ok\n");
  return 0;
}
