// Jump-ahead for the mt19937 stream (host side): what lets MANY CTAs produce one torch-exact word stream.
//
// The reference draws every mask and every shift tensor from torch's CPU generator (scheduler.py:279-291,431-448,
// 675,707), one serial recurrence x[k+624] = x[k+397] ^ twist(x[k], x[k+1]).  The recurrence is linear over GF(2) on
// the 19937-bit state, so the state J words ahead is g_J(T) s with g_J(x) = x^J mod phi(x), phi = the characteristic
// polynomial of the transition T (Haramoto, Matsumoto, Nishimura, Panneton, L'Ecuyer: "Efficient jump ahead for
// F2-linear random number generators", 2008).  Written out per word:  y[J + m] = XOR_{i : g_i = 1} y[i + m]  for any
// window of the untempered word sequence that starts at an index >= 1 (word 0 of a freshly seeded state carries 31
// bits that are not part of the state).  The device kernel (rng.cu: mt_stream_kernel) evaluates exactly that sum from
// 33 freshly generated blocks; this file computes phi (Berlekamp-Massey on one output bit), the polynomials
// g_c = x^{(c * blocks_per_cta - 1) * 624} mod phi for CTA c = 1 .. n, and a host utility that advances a state
// by n draws (used by the CPU tests to check the polynomial machinery against numpy's MT19937).
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace mdm {
namespace {

constexpr int MT_N = 624;
constexpr int DEG = 19937;
constexpr int PW = 312;   // 64-bit words of a polynomial / bit vector (19968 bits)

struct Poly {
  uint64_t w[PW];
  void clear() { memset(w, 0, sizeof(w)); }
  bool bit(int i) const { return (w[i >> 6] >> (i & 63)) & 1; }
  void flip(int i) { w[i >> 6] ^= 1ull << (i & 63); }
};

inline uint32_t twist(uint32_t u, uint32_t v) {
  return (((u & 0x80000000u) | (v & 0x7fffffffu)) >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}

// untempered word sequence y[0 .. count) continuing the block key[0 .. 624)
void mt_sequence(const uint32_t* key, uint32_t* y, int64_t count) {
  for (int64_t k = 0; k < count && k < MT_N; ++k) y[k] = key[k];
  for (int64_t k = MT_N; k < count; ++k) y[k] = y[k - 227] ^ twist(y[k - MT_N], y[k - MT_N + 1]);
}

inline void shl1(Poly& a) {
  uint64_t carry = 0;
  for (int i = 0; i < PW; ++i) {
    const uint64_t n = a.w[i] >> 63;
    a.w[i] = (a.w[i] << 1) | carry;
    carry = n;
  }
}
inline void xor_into(Poly& a, const Poly& b) {
  for (int i = 0; i < PW; ++i) a.w[i] ^= b.w[i];
}
// a ^= b << s (bits shifted beyond 19968 are dropped: callers keep degrees below that)
void xor_shifted(Poly& a, const Poly& b, int s) {
  const int ws = s >> 6, bs = s & 63;
  for (int i = PW - 1; i >= ws; --i) {
    uint64_t v = b.w[i - ws] << bs;
    if (bs && i - ws - 1 >= 0) v |= b.w[i - ws - 1] >> (64 - bs);
    a.w[i] ^= v;
  }
}

Poly g_phi;   // characteristic polynomial, degree 19937
std::once_flag g_phi_once;

// Berlekamp-Massey over GF(2) on bit 0 of y[1], y[2], ... of an arbitrary non-zero state: the transition's
// characteristic polynomial is primitive, so every non-trivial output bit sequence has it as minimal polynomial
void compute_phi() {
  const int NBITS = 2 * DEG + 64;
  std::vector<uint32_t> y(NBITS + 2);
  uint32_t key[MT_N];
  key[0] = 5489u;
  for (int j = 1; j < MT_N; ++j) key[j] = 1812433253u * (key[j - 1] ^ (key[j - 1] >> 30)) + (uint32_t)j;
  mt_sequence(key, y.data(), NBITS + 2);
  Poly C, B, R, T;
  C.clear(); B.clear(); R.clear();
  C.w[0] = 1; B.w[0] = 1;
  int L = 0, m = 1;
  for (int n = 0; n < NBITS; ++n) {
    const uint32_t s = y[n + 1] & 1u;
    // R bit i = s_{n-i} (i >= 1); discrepancy d = s_n ^ sum_{i=1..L} c_i s_{n-i}
    uint64_t acc = 0;
    const int words = (L >> 6) + 1;
    for (int i = 0; i < words; ++i) acc ^= C.w[i] & R.w[i];
    const uint32_t d = s ^ (uint32_t)(__builtin_popcountll(acc) & 1);
    if (d) {
      if (2 * L <= n) {
        T = C;
        xor_shifted(C, B, m);
        L = n + 1 - L;
        B = T;
        m = 1;
      } else {
        xor_shifted(C, B, m);
        ++m;
      }
    } else {
      ++m;
    }
    shl1(R);            // history for n + 1: bit i = s_{n+1-i}
    if (s) R.w[0] |= 2ull;
    R.w[0] &= ~1ull;
  }
  // connection polynomial C (s_n = sum_{i>=1} c_i s_{n-i}) -> characteristic polynomial phi_j = c_{L-j}
  g_phi.clear();
  if (L != DEG) return;   // leaves phi = 0: callers report the failure
  for (int j = 0; j <= L; ++j)
    if (C.bit(L - j)) g_phi.flip(j);
}

const Poly* phi() {
  std::call_once(g_phi_once, compute_phi);
  return g_phi.bit(DEG) ? &g_phi : nullptr;
}

inline void mul_x(Poly& a, const Poly& ph) {
  shl1(a);
  if (a.bit(DEG)) xor_into(a, ph);
}
// a * b mod phi (degrees < 19937)
Poly mulmod(const Poly& a, const Poly& b, const Poly& ph) {
  Poly acc;
  acc.clear();
  for (int i = DEG - 1; i >= 0; --i) {
    mul_x(acc, ph);
    if (b.bit(i)) xor_into(acc, a);
  }
  return acc;
}
// x^e mod phi
Poly xpow(uint64_t e, const Poly& ph) {
  Poly r;
  r.clear();
  r.w[0] = 1;
  int top = 63;
  while (top >= 0 && !((e >> top) & 1)) --top;
  for (int b = top; b >= 0; --b) {
    r = mulmod(r, r, ph);
    if ((e >> b) & 1) mul_x(r, ph);
  }
  return r;
}

// window[m] = XOR_{i : g_i} seq[i + m], m = 0 .. 623 (seq holds >= 19937 + 623 words)
void apply_poly(const Poly& g, const uint32_t* seq, uint32_t* window) {
  memset(window, 0, MT_N * sizeof(uint32_t));
  for (int i = 0; i < DEG; ++i)
    if (g.bit(i))
      for (int m = 0; m < MT_N; ++m) window[m] ^= seq[i + m];
}

}  // namespace
}  // namespace mdm

using namespace mdm;

extern "C" {

// polys_host[(c - 1) * 624 .. c * 624) = bits of x^{(c * blocks_per_cta - 1) * 624} mod phi, c = 1 .. n_polys
// (bit i of the polynomial = bit (i & 31) of word i >> 5).  Host threads share the work.
int mdm_rng_jump_table_host(uint32_t* polys_host, int n_polys, int blocks_per_cta) {
  MDM_CHECK_ARG(polys_host && n_polys >= 1 && blocks_per_cta >= 2, "jump_table: bad arguments");
  const Poly* ph = phi();
  if (!ph) { set_error("jump_table: Berlekamp-Massey did not find a degree-19937 polynomial"); return MDM_E_UNSUPPORTED; }
  int T = (int)std::thread::hardware_concurrency();
  if (T < 1) T = 1;
  if (T > 32) T = 32;
  if (T > n_polys) T = n_polys;
  const uint64_t stride_words = (uint64_t)blocks_per_cta * MT_N;
  const Poly step = xpow(stride_words * (uint64_t)T, *ph);   // x^{T * W}
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t) {
    th.emplace_back([=, &step]() {
      Poly g = xpow(((uint64_t)(t + 1) * blocks_per_cta - 1) * MT_N, *ph);
      for (int c = t + 1; c <= n_polys; c += T) {
        memcpy(polys_host + (size_t)(c - 1) * MT_N, g.w, MT_N * sizeof(uint32_t));
        if (c + T <= n_polys) g = mulmod(g, step, *ph);
      }
    });
  }
  for (auto& x : th) x.join();
  return MDM_OK;
}

// host utility (no device work): the state torch's CPU generator holds after n more draws from state_in
// (625 words: key + position).  Far jumps use the polynomial method -- the same arithmetic the device kernel uses.
int mdm_rng_advance_host(const uint32_t* state_in, int64_t n, uint32_t* state_out) {
  MDM_CHECK_ARG(state_in && state_out && n >= 0, "advance_host: bad arguments");
  const uint32_t pos = state_in[MT_N];
  MDM_CHECK_ARG(pos <= MT_N, "advance_host: position %u out of range", pos);
  const int64_t P = (int64_t)pos + n;
  const int64_t b = P == 0 ? 0 : (P + MT_N - 1) / MT_N - 1;   // torch block that holds word P - 1
  if (b == 0) {
    memcpy(state_out, state_in, (MT_N + 1) * sizeof(uint32_t));
    state_out[MT_N] = (uint32_t)P;
    return MDM_OK;
  }
  std::vector<uint32_t> y;
  if (b <= 40) {
    y.resize((size_t)(b + 1) * MT_N);
    mt_sequence(state_in, y.data(), (int64_t)y.size());
    memcpy(state_out, y.data() + b * MT_N, MT_N * sizeof(uint32_t));
  } else {
    const Poly* ph = phi();
    if (!ph) { set_error("advance_host: characteristic polynomial unavailable"); return MDM_E_UNSUPPORTED; }
    // z[k] = y[k + 1]; window at z-offset J = (b - 1) * 624 = y[624 (b-1) + 1 .. 624 b], then one more block
    y.resize(1 + DEG + MT_N + 8);
    mt_sequence(state_in, y.data(), (int64_t)y.size());
    const Poly g = xpow((uint64_t)(b - 1) * MT_N, *ph);
    uint32_t win[2 * MT_N];
    apply_poly(g, y.data() + 1, win);
    for (int k = 0; k < MT_N; ++k) win[MT_N + k] = win[MT_N + k - 227] ^ twist(win[k], win[k + 1]);
    state_out[0] = win[MT_N - 1];                                  // y[624 b]
    memcpy(state_out + 1, win + MT_N, (MT_N - 1) * sizeof(uint32_t));   // y[624 b + 1 .. 624 b + 623]
  }
  state_out[MT_N] = (uint32_t)(P - b * MT_N);
  return MDM_OK;
}

}  // extern "C"
