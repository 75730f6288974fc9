// zkstark_b200.hpp -- C++17 host-side mirror of the reference crate's interface for the hot path, over the C ABI
// in zkstark_b200.h.  The reference (Crocodoctopus/zkstark) is a Rust crate; its toolchain is absent from the build
// image, so the host layer that a Rust shim would provide (bindings/rust/) is provided in C++ with the same
// names, argument meaning and error behaviour:
//
//   reference (Rust)                                         here (namespace zk)
//   Merkle::new(size, iter) / merkle[i] / merkle.trace(i)    Merkle::create(size, data) / m[i] / m.trace(i)   merkle.rs:14,54,74
//   compute_root_from_path(element, index, &path)            compute_root_from_path(...)                      merkle.rs:82
//   lagrange::<F>(&points)  (x_i = g^i)                      lagrange(points) / lagrange_subgroup(ys, log_n)  polynomial.rs:337
//   f_poly.solve(x) over f_domain                            solve_on_coset(coeffs, log_n)                    prover.rs:69-70
//   fri::<F>(&poly, beta) + solve on the squared domain      fri_fold(cur, layer_k, beta_raw)                 polynomial.rs:385
//   Channel::new() / commit / get_u32 / finalize             Channel{} / commit(..) / get_u32() / finalize()  channel.rs:11-37
//   generate_proof(channel) -> Proof                         generate_proof(channel, log_n)                   prover.rs:9
//   proof.verify() (panics) / proof.size()                   proof.verify() (throws) / proof.size()           proof.rs:15,151
//
// Errors: the reference panics (assert_eq!/unwrap); every failure here throws zk::Error.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "zkstark_b200.h"

namespace zk {

using Hash = std::array<uint8_t, 32>;   // merkle.rs:9
using AuthPath = std::vector<Hash>;     // merkle.rs:10
constexpr uint32_t MODULUS = ZK_MODULUS;  // main.rs:13

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// One zk_ctx per thread (the C ABI is not thread-safe), created on first use.  No CPU fallback.
class Context {
public:
    static Context& current()
    {
        thread_local Context ctx;
        return ctx;
    }
    zk_ctx* get() const { return ctx_; }
    void check(int rc) const
    {
        if (rc != ZK_OK) throw Error(rc, std::string("zkstark_b200: ") + zk_last_error(ctx_));
    }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    ~Context() { zk_ctx_destroy(ctx_); }

private:
    Context()
    {
        int rc = zk_ctx_create(0, &ctx_);
        if (rc != ZK_OK) throw Error(rc, "zk_ctx_create failed: no CUDA device (there is no CPU fallback)");
    }
    zk_ctx* ctx_ = nullptr;
};

// ---------------------------------------------------------------- merkle.rs
class Merkle {
public:
    // Merkle::new(size, data)  (merkle.rs:14-51)
    static Merkle create(size_t size, const std::vector<uint32_t>& data)
    {
        if (size == 0 || (size & (size - 1))) throw Error(ZK_ERR_INVALID, "size must be a power of two (merkle.rs:17-21)");
        if (data.size() != size) throw Error(ZK_ERR_INVALID, "iterator length must equal size (merkle.rs:28-35)");
        Merkle m;
        m.heap_.resize(2 * size - 1);
        auto& c = Context::current();
        c.check(zk_merkle_build(c.get(), data.data(), size, reinterpret_cast<uint8_t*>(m.heap_.data())));
        return m;
    }
    const Hash& operator[](size_t i) const { return heap_.at(i); }  // merkle.rs:74-79
    AuthPath trace(size_t i) const                                  // merkle.rs:54-71
    {
        const size_t size = (heap_.size() + 1) / 2;
        AuthPath p(64);
        int n = zk_merkle_trace(reinterpret_cast<const uint8_t*>(heap_.data()), size, i, reinterpret_cast<uint8_t*>(p.data()));
        if (n < 0) throw Error(n, "leaf index out of range");
        p.resize(n);
        return p;
    }
    size_t len() const { return heap_.size(); }

private:
    std::vector<Hash> heap_;
};

// ---------------------------------------------------------------- polynomial.rs (hot-path shapes)
// lagrange(&points) for points (g^i, ys[i]); ys.size() is n-1 (prover.rs:60: zip truncates) or n.  Ascending coefficients.
inline std::vector<uint32_t> lagrange_subgroup(const std::vector<uint32_t>& ys, uint32_t log_n)
{
    std::vector<uint32_t> c(size_t(1) << log_n);
    auto& x = Context::current();
    x.check(zk_interpolate_subgroup(x.get(), ys.data(), static_cast<uint32_t>(ys.size()), log_n, c.data()));
    return c;
}
// f_domain.iter().map(|&n| f_poly.solve(n))  (prover.rs:69-70): evaluations at 5 h^i, natural order
inline std::vector<uint32_t> solve_on_coset(const std::vector<uint32_t>& coeffs, uint32_t log_n, uint32_t shift = 5)
{
    std::vector<uint32_t> e(size_t(8) << log_n);
    auto& x = Context::current();
    x.check(zk_lde_coset(x.get(), coeffs.data(), log_n, ZK_LOG_BLOWUP, shift, e.data()));
    return e;
}
// fri(poly, beta) then evaluation on the squared half domain (polynomial.rs:385-400, prover.rs:200-211)
inline std::vector<uint32_t> fri_fold(const std::vector<uint32_t>& cur, uint32_t layer_k, uint32_t beta_raw)
{
    std::vector<uint32_t> nx(cur.size() / 2);
    auto& x = Context::current();
    x.check(zk_fri_fold(x.get(), cur.data(), cur.size(), layer_k, beta_raw, nx.data()));
    return nx;
}
// the three constraints + alpha mix evaluated on the LDE domain (prover.rs:101-173)
inline std::vector<uint32_t> cp_eval_fibsq(const std::vector<uint32_t>& f_eval, uint32_t log_n, const uint32_t alpha_raw[3],
                                           uint32_t a_first, uint32_t a_last)
{
    std::vector<uint32_t> cp(f_eval.size());
    auto& x = Context::current();
    x.check(zk_cp_eval_fibsq(x.get(), f_eval.data(), log_n, ZK_LOG_BLOWUP, alpha_raw, a_first, a_last, cp.data()));
    return cp;
}
// prover.rs:32-39
inline std::vector<uint32_t> trace_fibsq(uint32_t log_n, uint32_t secret)
{
    std::vector<uint32_t> a((size_t(1) << log_n) - 1);
    if (zk_trace_fibsq(log_n, secret, a.data()) != ZK_OK) throw Error(ZK_ERR_INVALID, "bad log_n");
    return a;
}

// lagrange::<F>(&points) with the reference's own signature (polynomial.rs:337): points (x_i, y_i).  The GPU path
// interpolates over the trace subgroup, so the abscissae must be x_i = g^i, g = 5^((p-1)/n), n the next power of two
// >= points.size() -- what prover.rs:56-61 builds; anything else throws (there is no CPU fallback).  Ascending
// coefficients without trailing zeros, i.e. `Polynomial.0` (polynomial.rs:15-28,31).
inline std::vector<uint32_t> lagrange(const std::vector<std::pair<uint32_t, uint32_t>>& points)
{
    size_t n = 1;
    uint32_t log_n = 0;
    while (n < points.size()) { n <<= 1; log_n++; }
    if (points.size() != n && points.size() != n - 1)
        throw Error(ZK_ERR_INVALID, "lagrange on the GPU needs n or n - 1 points of the 2^k subgroup");
    uint64_t g = 1, b = 5, e = (uint64_t(MODULUS) - 1) >> log_n;   // g = 5^((p-1)/n)  (prover.rs:48)
    for (; e; e >>= 1, b = b * b % MODULUS)
        if (e & 1) g = g * b % MODULUS;
    std::vector<uint32_t> ys(points.size());
    uint64_t x = 1;
    for (size_t i = 0; i < points.size(); i++, x = x * g % MODULUS) {
        if (points[i].first != x) throw Error(ZK_ERR_INVALID, "lagrange on the GPU needs abscissae g^i (prover.rs:56-61)");
        ys[i] = points[i].second;
    }
    std::vector<uint32_t> c = lagrange_subgroup(ys, log_n);
    while (!c.empty() && c.back() == 0) c.pop_back();
    return c;
}

// ---------------------------------------------------------------- channel.rs / proof.rs / prover.rs
struct Proof;
// channel.rs:6-37 with bincode 1.x's layout of what prover.rs commits (u32: 4 bytes LE; [u8; 32]: raw; Box<[Hash]>: u64 LE
// length + digests; tuples: concatenation).  SHA-256 is the library's host implementation (sha2 in the reference).
struct Channel {
    Hash state{};                 // channel.rs:12-17: Channel::new() is the zero state, empty transcript
    std::vector<uint8_t> data;

    void commit_bytes(const std::vector<uint8_t>& v)   // channel.rs:19-26
    {
        std::vector<uint8_t> msg(state.begin(), state.end());
        msg.insert(msg.end(), v.begin(), v.end());
        zk_sha256(msg.data(), msg.size(), state.data());
        data.insert(data.end(), v.begin(), v.end());
    }
    void commit(const Hash& h) { commit_bytes(std::vector<uint8_t>(h.begin(), h.end())); }
    void commit(uint32_t v) { std::vector<uint8_t> b; put_u32(b, v); commit_bytes(b); }
    void commit(uint32_t v, const AuthPath& p)         // prover.rs:274-277: (value, path)
    {
        std::vector<uint8_t> b;
        put_u32(b, v);
        put_path(b, p);
        commit_bytes(b);
    }
    void commit(uint32_t a, uint32_t b2, const AuthPath& p, const AuthPath& q)   // prover.rs:288: (cp_x, cp_nx, path, path)
    {
        std::vector<uint8_t> b;
        put_u32(b, a);
        put_u32(b, b2);
        put_path(b, p);
        put_path(b, q);
        commit_bytes(b);
    }
    uint32_t get_u32()                                 // channel.rs:28-32: big-endian read, un-reduced, committed
    {
        const uint32_t f = (uint32_t(state[0]) << 24) | (uint32_t(state[1]) << 16) | (uint32_t(state[2]) << 8) | state[3];
        commit(f);
        return f;
    }
    inline Proof finalize() const;                     // channel.rs:34-36

private:
    static void put_u32(std::vector<uint8_t>& b, uint32_t v)
    {
        for (int i = 0; i < 4; i++) b.push_back(uint8_t(v >> (8 * i)));
    }
    static void put_path(std::vector<uint8_t>& b, const AuthPath& p)
    {
        const uint64_t n = p.size();
        for (int i = 0; i < 8; i++) b.push_back(uint8_t(n >> (8 * i)));
        for (const Hash& h : p) b.insert(b.end(), h.begin(), h.end());
    }
};

struct Proof {  // proof.rs:5-13
    Hash state{};
    std::vector<uint8_t> data;
    uint32_t log_n = 10;
    uint32_t a_last = 2338775057u;  // proof.rs:70
    zk_air air{1, 1, 1, 0, 0, 0, 0, 0};  // the program the proof is about; the reference's FibonacciSq
    // proof.rs:15-149; throws where the reference panics.  Also replays the channel (alpha/beta/query, state).
    void verify() const
    {
        int line = 0;
        int rc = zk_verify_air(&air, log_n, a_last, data.data(), data.size(), state.data(), nullptr, 0, 1, &line);
        if (rc != ZK_OK) throw Error(rc, "proof rejected: the assert at proof.rs:" + std::to_string(line) + " fails");
    }
    size_t size() const { return 48 + data.size(); }  // proof.rs:151-154 (size_of::<Proof>() == 48)
    // wire format: what bincode 1.x writes for the struct (32 state bytes, u64 LE length, data)
    std::vector<uint8_t> to_bytes() const
    {
        std::vector<uint8_t> out(zk_proof_serialized_len(data.size()));
        size_t n = 0;
        int rc = zk_proof_serialize(state.data(), data.data(), data.size(), out.data(), out.size(), &n);
        if (rc != ZK_OK) throw Error(rc, "zk_proof_serialize");
        out.resize(n);
        return out;
    }
    static Proof from_bytes(const std::vector<uint8_t>& blob, uint32_t log_n = 10, uint32_t a_last = 2338775057u)
    {
        Proof p;
        const uint8_t* tr = nullptr;
        size_t n = 0;
        int rc = zk_proof_deserialize(blob.data(), blob.size(), p.state.data(), &tr, &n);
        if (rc != ZK_OK) throw Error(rc, rc == ZK_ERR_BUFFER ? "truncated proof" : "proof length prefix does not match the payload");
        p.data.assign(tr, tr + n);
        p.log_n = log_n;
        p.a_last = a_last;
        return p;
    }
};
inline Proof Channel::finalize() const
{
    Proof p;
    p.state = state;
    p.data = data;
    return p;
}

// generate_proof(channel) (prover.rs:9-293), generalised to n = 2^log_n; the reference is log_n = 10, secret 3141592.
inline Proof generate_proof(const Channel& channel, uint32_t log_n = 10, uint32_t secret = 3141592u)
{
    if (!channel.data.empty() || channel.state != Hash{})
        throw Error(ZK_ERR_INVALID, "generate_proof expects Channel::new() (channel.rs:12-17)");
    Proof p;
    p.log_n = log_n;
    p.data.resize(zk_transcript_len(log_n));
    size_t len = 0;
    auto& x = Context::current();
    x.check(zk_prove_fibsq(x.get(), log_n, secret, p.data.data(), p.data.size(), &len, p.state.data()));
    p.data.resize(len);
    p.a_last = trace_fibsq(log_n, secret).back();
    return p;
}

// generate_proof for another program of the zk_air family (the context proves `air` for the duration of the
// call and returns to the reference's program afterwards).
inline std::vector<uint32_t> trace_air(const zk_air& air, uint32_t log_n, uint32_t secret)
{
    std::vector<uint32_t> a((size_t(1) << log_n) - 1);
    if (zk_trace_air(&air, log_n, secret, a.data()) != ZK_OK) throw Error(ZK_ERR_INVALID, "bad log_n or zk_air");
    return a;
}
inline Proof generate_proof_air(const Channel& channel, const zk_air& air, uint32_t log_n, uint32_t secret)
{
    auto& x = Context::current();
    x.check(zk_ctx_set_air(x.get(), &air));
    struct Reset {
        zk_ctx* c;
        ~Reset() { zk_ctx_set_air(c, nullptr); }
    } reset{x.get()};
    Proof p = generate_proof(channel, log_n, secret);
    p.air = air;
    p.a_last = trace_air(air, log_n, secret).back();
    return p;
}

}  // namespace zk
