// prover.rs:31-292 replayed CALL BY CALL through the per-function exports of the C ABI (via the C++ mirror
// include/zkstark_b200.hpp), with the Fiat-Shamir channel on the HOST -- the integration a maintainer gets by keeping
// prover.rs and swapping the bodies it calls (INTEGRATION.md "per function"):
//   lagrange -> solve over the coset -> Merkle::new -> channel.commit -> constraints -> Merkle::new -> (fri + solve +
//   Merkle::new + commit) x log_n -> free term -> query -> merkle.trace x (4 + 2 log_n) -> finalize.
// The transcript must equal the monolithic device-resident pipeline's (zk_prove_fibsq) byte for byte, and its
// sha256 the golden value given on the command line (tests/golden/transcripts.json: b956f693... for the reference's
// own configuration, f404e6df... for 2^16).
//   usage: prover_replay <log_n> <sha256 hex of the transcript> [<log_n> <sha256> ...]
#include <cstdio>
#include <cstring>
#include <string>

#include "zkstark_b200.hpp"

using namespace zk;

static std::string hex(const uint8_t* p, size_t n)
{
    static const char* d = "0123456789abcdef";
    std::string s;
    for (size_t i = 0; i < n; i++) { s += d[p[i] >> 4]; s += d[p[i] & 15]; }
    return s;
}

static uint64_t pow_mod(uint64_t b, uint64_t e)
{
    uint64_t r = 1;
    for (b %= MODULUS; e; e >>= 1, b = b * b % MODULUS)
        if (e & 1) r = r * b % MODULUS;
    return r;
}

// generate_proof (prover.rs:9-293) for n = 2^log_n with the reference's call structure
static Proof generate_proof_per_function(Channel channel, uint32_t log_n, uint32_t secret)
{
    const size_t n = size_t(1) << log_n, N = 8 * n;
    // Part 1 (prover.rs:31-85)
    const std::vector<uint32_t> a = trace_fibsq(log_n, secret);                        // :32-39
    const uint64_t generator_g = pow_mod(5, (uint64_t(MODULUS) - 1) >> log_n);         // :48
    std::vector<std::pair<uint32_t, uint32_t>> points;                                 // :60  zip(&g, &a): n - 1 pairs
    uint64_t gi = 1;
    for (size_t i = 0; i + 1 < n; i++, gi = gi * generator_g % MODULUS) points.emplace_back(uint32_t(gi), a[i]);
    std::vector<uint32_t> f_poly = lagrange(points);                                   // :61
    if (f_poly.size() != n - 1) throw Error(ZK_ERR_ASSERT, "f_poly must have degree n - 2");
    f_poly.resize(n, 0);
    const std::vector<uint32_t> f_eval = solve_on_coset(f_poly, log_n);                // :69-70
    const Merkle f_eval_merkle = Merkle::create(N, f_eval);                            // :81
    channel.commit(f_eval_merkle[0]);                                                  // :82-85
    // Part 2 (prover.rs:101-180)
    uint32_t alpha_raw[3];
    for (int j = 0; j < 3; j++) alpha_raw[j] = channel.get_u32();                      // :163-165
    std::vector<std::vector<uint32_t>> cp_evals;
    std::vector<Merkle> cp_eval_merkles;
    cp_evals.push_back(cp_eval_fibsq(f_eval, log_n, alpha_raw, a[0], a[n - 2]));        // :101-173
    cp_eval_merkles.push_back(Merkle::create(N, cp_evals[0]));                         // :176
    channel.commit(cp_eval_merkles[0][0]);                                             // :177-180
    // Part 3 (prover.rs:192-254)
    for (uint32_t i = 0; i < log_n; i++) {                                             // :198  `for _ in 0..10`
        const uint32_t beta_raw = channel.get_u32();                                   // :200
        cp_evals.push_back(fri_fold(cp_evals.back(), i, beta_raw));                    // :201-211
        cp_eval_merkles.push_back(Merkle::create(cp_evals.back().size(), cp_evals.back()));   // :214
        channel.commit(cp_eval_merkles.back()[0]);                                     // :224
    }
    for (uint32_t v : cp_evals.back())                                                 // :238  degree 0: a constant layer
        if (v != cp_evals.back()[0]) throw Error(ZK_ERR_ASSERT, "last FRI layer is not constant (prover.rs:238)");
    if (cp_evals.back().size() != 8) throw Error(ZK_ERR_ASSERT, "last FRI layer must have 8 values (prover.rs:251)");
    channel.commit(cp_evals.back()[0]);                                                // :254  the free term
    // Part 4 (prover.rs:263-292)
    const size_t x = size_t(channel.get_u32()) % (N - 16);                             // :263
    channel.commit(f_eval[x], f_eval_merkle.trace(x));                                 // :266-274
    channel.commit(f_eval[x + 8], f_eval_merkle.trace(x + 8));                         // :268-275
    channel.commit(f_eval[x + 16], f_eval_merkle.trace(x + 16));                       // :270-276
    channel.commit(cp_evals[0][x], cp_eval_merkles[0].trace(x));                       // :272-277
    for (uint32_t i = 0; i < log_n; i++) {                                             // :280-289
        const size_t len = cp_evals[i].size();
        const size_t xi = x % len, nx = (xi + len / 2) % len;
        channel.commit(cp_evals[i][xi], cp_evals[i][nx], cp_eval_merkles[i].trace(xi), cp_eval_merkles[i].trace(nx));
    }
    Proof p = channel.finalize();                                                      // :292
    p.log_n = log_n;
    p.a_last = a[n - 2];
    return p;
}

int main(int argc, char** argv)
{
    try {
        for (int i = 1; i + 1 < argc; i += 2) {
            const uint32_t log_n = uint32_t(std::stoul(argv[i]));
            const std::string want = argv[i + 1];
            const Proof p = generate_proof_per_function(Channel{}, log_n, 3141592u);
            Hash d;
            zk_sha256(p.data.data(), p.data.size(), d.data());
            const Proof q = generate_proof(Channel{}, log_n, 3141592u);                // the monolithic pipeline
            std::printf("log_n %u: per-function transcript %zu bytes sha256 %s state %s\n", log_n, p.data.size(),
                        hex(d.data(), 32).c_str(), hex(p.state.data(), 32).c_str());
            if (hex(d.data(), 32) != want) { std::printf("FAIL: golden is %s\n", want.c_str()); return 1; }
            if (p.data != q.data || p.state != q.state) { std::printf("FAIL: differs from zk_prove_fibsq\n"); return 1; }
            p.verify();                                                                // proof.rs:15-149 + channel replay
            // wire format round trip through the C ABI
            const std::vector<uint8_t> blob = p.to_bytes();
            const Proof back = Proof::from_bytes(blob, log_n, p.a_last);
            if (blob.size() != 40 + p.data.size() || back.data != p.data || back.state != p.state) { std::printf("FAIL: wire format\n"); return 1; }
            back.verify();
            bool threw = false;
            try { Proof::from_bytes(std::vector<uint8_t>(blob.begin(), blob.end() - 1)); } catch (const Error&) { threw = true; }
            if (!threw) { std::printf("FAIL: truncated proof accepted\n"); return 1; }
        }
        // another program of the AIR family through the same entry point; its proof verifies under its own
        // program and under no other
        {
            const zk_air fib{1, 0, 0, 0, 1, 1, 0, 0};                                     // a[i+2] = a[i+1] + a[i]
            Proof p = generate_proof_air(Channel{}, fib, 12, 2u);
            p.verify();
            if (p.a_last != trace_air(fib, 12, 2u).back()) { std::printf("FAIL: air a_last\n"); return 1; }
            bool rejected = false;
            p.air = zk_air{1, 1, 1, 0, 0, 0, 0, 0};
            try { p.verify(); } catch (const Error&) { rejected = true; }
            if (!rejected) { std::printf("FAIL: a Fibonacci proof verified as FibonacciSq\n"); return 1; }
            const Proof q = generate_proof(Channel{}, 10, 3141592u);                      // the context is back on FibonacciSq
            q.verify();
        }
        // lagrange() refuses abscissae that are not g^i: there is no CPU fallback to fall into
        bool threw = false;
        try { lagrange({{1u, 5u}, {7u, 9u}, {11u, 2u}}); } catch (const Error&) { threw = true; }
        if (!threw) { std::printf("FAIL: lagrange accepted arbitrary abscissae\n"); return 1; }
    } catch (const Error& e) {
        std::printf("FAIL: %s (code %d)\n", e.what(), e.code);
        return 1;
    }
    std::printf("prover replay ok\n");
    return 0;
}
