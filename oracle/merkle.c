/*
 * oracle/merkle.c -- CPU ORACLE (test infrastructure).  Restates merkle.rs.
 */
#include "oracle.h"
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_sha256_blocks(uint32_t st[8], const uint8_t *data, size_t nblocks);

static const uint32_t IV[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                               0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
static void store_digest(const uint32_t st[8], uint8_t *out)
{
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(st[i] >> 24); out[4 * i + 1] = (uint8_t)(st[i] >> 16);
        out[4 * i + 2] = (uint8_t)(st[i] >> 8); out[4 * i + 3] = (uint8_t)st[i];
    }
}
/* SHA256 of a 4-byte message: one padded block (same digest as orc_sha256(be, 4)) */
static void hash_leaf(uint32_t v, uint8_t *out)
{
    uint8_t blk[64] = {0};
    uint32_t st[8];
    memcpy(st, IV, sizeof st);
    blk[0] = (uint8_t)(v >> 24); blk[1] = (uint8_t)(v >> 16); blk[2] = (uint8_t)(v >> 8); blk[3] = (uint8_t)v;
    blk[4] = 0x80; blk[63] = 32;
    orc_sha256_blocks(st, blk, 1);
    store_digest(st, out);
}
/* SHA256 of a 64-byte message: the data block, then the constant padding block */
static void hash_node(const uint8_t *lr, uint8_t *out)
{
    static const uint8_t pad[64] = {0x80, [62] = 0x02, [63] = 0x00};
    uint32_t st[8];
    memcpy(st, IV, sizeof st);
    orc_sha256_blocks(st, lr, 1);
    orc_sha256_blocks(st, pad, 1);
    store_digest(st, out);
}

/* merkle.rs:14-51 */
int orc_merkle_build(const uint32_t *leaves, uint64_t size, uint8_t *heap, int threads)
{
    if (size == 0 || (size & (size - 1))) return -1; /* merkle.rs:17-21 assert power of two */
    (void)threads;
    uint64_t offset = size - 1;                      /* merkle.rs:27 out.len()/2 */
    /* merkle.rs:28-35: leaf = SHA256(v.to_be_bytes()) */
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1)
    for (int64_t i = 0; i < (int64_t)size; i++) {
        hash_leaf(leaves[i], heap + 32 * (offset + i));
    }
    /* merkle.rs:38-47: node = SHA256(left || right), children of i at 2i+1, 2i+2 */
    while (offset > 0) {
        offset /= 2;
        int64_t cnt = (int64_t)offset + 1;
#pragma omp parallel for schedule(static) num_threads(threads > 0 ? threads : 1) if (cnt >= 1024)
        for (int64_t it = 0; it < cnt; it++) {
            uint64_t idx = offset + it;
            hash_node(heap + 32 * (2 * idx + 1), heap + 32 * idx);
        }
    }
    return 0;
}

/* merkle.rs:54-71 */
int orc_merkle_trace(const uint8_t *heap, uint64_t size, uint64_t i, uint8_t *path_out)
{
    int n = 0;
    i += size - 1; /* self.0.len()/2 */
    while (i != 0) {
        if (i % 2 == 0) { /* right node, store left */
            memcpy(path_out + 32 * n, heap + 32 * (i - 1), 32);
            i -= 2;
        } else {          /* left node, store right */
            memcpy(path_out + 32 * n, heap + 32 * (i + 1), 32);
            i -= 1;
        }
        i >>= 1;
        n++;
    }
    return n;
}

/* merkle.rs:82-110 */
void orc_root_from_path(uint32_t element, uint64_t index, const uint8_t *path, uint32_t path_len,
                        uint8_t out[32])
{
    uint8_t cur[32], buf[64];
    index += ((uint64_t)1 << path_len) - 1;
    uint8_t be[4] = {(uint8_t)(element >> 24), (uint8_t)(element >> 16), (uint8_t)(element >> 8),
                     (uint8_t)element};
    orc_sha256(be, 4, cur);
    for (uint32_t k = 0; k < path_len; k++) {
        const uint8_t *h = path + 32 * k;
        if (index % 2 == 0) {
            memcpy(buf, h, 32); memcpy(buf + 32, cur, 32);
            index -= 2;
        } else {
            memcpy(buf, cur, 32); memcpy(buf + 32, h, 32);
            index -= 1;
        }
        orc_sha256(buf, 64, cur);
        index >>= 1;
    }
    memcpy(out, cur, 32);
}
