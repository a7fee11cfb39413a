// host_twin.cpp -- TEST INFRASTRUCTURE: compiles the product's bit logic
// (ecen743-project-cgol_b200/csrc/cgl_bits.cuh, the very expressions the sm_100a kernels run)
// with g++ so the CPU box can check it against the oracle before any GPU time is spent.
// Loop structure mirrors the kernels in cgl_env.cu (generic and fused paths); it is not shipped.
#include <stdint.h>
#include <string.h>
#include <vector>

#include "../../ecen743-project-cgol_b200/csrc/cgl_bits.cuh"

using namespace cgl;

extern "C" {

void twin_pack(const uint8_t *cells, uint32_t *world, uint64_t n_envs, uint32_t rows, uint32_t cols)
{
    const uint32_t W = (cols + 31) / 32;
    for (uint64_t row = 0; row < n_envs * rows; ++row)
        for (uint32_t w = 0; w < W; ++w) {
            uint32_t word = 0;
            for (uint32_t j = 0; j < 32; ++j) {
                uint32_t col = w * 32 + j;
                if (col < cols && cells[row * cols + col] != 0) word |= 1u << j;
            }
            world[row * W + w] = word;
        }
}

void twin_unpack(const uint32_t *world, uint8_t *cells, uint64_t n_envs, uint32_t rows, uint32_t cols)
{
    const uint32_t W = (cols + 31) / 32;
    for (uint64_t row = 0; row < n_envs * rows; ++row)
        for (uint32_t c = 0; c < cols; ++c) cells[row * cols + c] = (world[row * W + (c >> 5)] >> (c & 31)) & 1u;
}

// mirrors life_generic_kernel
void twin_life_generic(const uint32_t *in, uint32_t *out, uint64_t n_envs, uint32_t rows, uint32_t cols,
                       int wrap_rows)
{
    const uint32_t W = (cols + 31) / 32;
    const uint64_t wpe = (uint64_t)rows * W;
    const uint32_t rbits = cols - 32 * (W - 1);
    const uint32_t last_mask = rbits == 32 ? 0xffffffffu : ((1u << rbits) - 1u);
    for (uint64_t e = 0; e < n_envs; ++e)
        for (uint32_t r = 0; r < rows; ++r)
            for (uint32_t w = 0; w < W; ++w) {
                const uint32_t *base = in + e * wpe;
                RowPlanes pc = load_row_planes(base + (uint64_t)r * W, w, W, rbits);
                RowPlanes pa = {0, 0, 0}, pb = {0, 0, 0};
                if (r > 0) pa = load_row_planes(base + (uint64_t)(r - 1) * W, w, W, rbits);
                else if (wrap_rows) pa = load_row_planes(base + (uint64_t)(rows - 1) * W, w, W, rbits);
                if (r + 1 < rows) pb = load_row_planes(base + (uint64_t)(r + 1) * W, w, W, rbits);
                else if (wrap_rows) pb = load_row_planes(base, w, W, rbits);
                uint32_t nxt = life_rule(hsum(pa.west, pa.c, pa.east), hsum(pc.west, pc.c, pc.east),
                                         hsum(pb.west, pb.c, pb.east), pc.c);
                if (w == W - 1) nxt &= last_mask;
                out[e * wpe + (uint64_t)r * W + w] = nxt;
            }
}

// mirrors stable_generic_kernel
void twin_stable_generic(const uint32_t *prev, const uint32_t *next, int8_t *stable, uint64_t n_envs,
                         uint32_t side, int spawn, int stable_max)
{
    const uint32_t W = (side + 31) / 32;
    const uint64_t size = (uint64_t)side * side;
    for (uint64_t e = 0; e < n_envs; ++e)
        for (uint32_t r = 0; r < side; ++r)
            for (uint32_t c = 0; c < side; ++c) {
                const uint64_t widx = (e * side + r) * W + (c >> 5);
                const bool p = (prev[widx] >> (c & 31)) & 1u, n = (next[widx] >> (c & 31)) & 1u;
                int8_t &s = stable[e * size + (uint64_t)r * side + c];
                s = stable_update1(s, p, n, (int8_t)spawn, (int8_t)stable_max);
            }
}

// mirrors env_step_fused_kernel (side % 32 == 0): phases A (toggle), B (rule + nibble mix), C (LUT + SIMD)
int twin_env_step_fused(const uint32_t *world_in, uint32_t *world_out, int8_t *stable, uint64_t n_envs,
                        uint32_t side, const int32_t *actions, int spawn, int stable_max, int32_t *reward_out,
                        uint32_t *alive_out)
{
    if (side % 32) return -1;
    const int S = (int)side, W = S / 32, WPE = S * W, SIZE = S * S, NCHUNK = SIZE / 16;
    const uint32_t spawn4 = rep4(spawn), max4 = rep4(stable_max);
    uint32_t table[16][2];                       // [nibble][0] = byte mask, [1] = mask & SPAWN
    for (uint32_t i = 0; i < 16; ++i) { table[i][0] = nibble_to_bytemask(i); table[i][1] = table[i][0] & spawn4; }
    std::vector<uint32_t> cur(WPE), mix(2 * WPE);
    int err = 0;
    for (uint64_t e = 0; e < n_envs; ++e) {
        int act = -1;
        if (actions) {
            int a = actions[e];
            if (a >= 0 && a < SIZE) act = a;
            else if (a != SIZE) err = 1;
        }
        memcpy(cur.data(), world_in + e * WPE, WPE * 4);
        if (act >= 0) cur[act >> 5] ^= 1u << (act & 31);
        uint32_t pop = 0;
        for (int i = 0; i < WPE; ++i) {
            const int r = i / W, w = i % W;
            const int ru = (r == 0 ? S - 1 : r - 1) * W, rc = r * W, rd = (r == S - 1 ? 0 : r + 1) * W;
            const int wl = (w == 0 ? W - 1 : w - 1), wr = (w == W - 1 ? 0 : w + 1);
            const uint32_t a = cur[ru + w], c = cur[rc + w], b = cur[rd + w];
            const HSum ha = hsum(west_plane(cur[ru + wl], a), a, east_plane(a, cur[ru + wr]));
            const HSum hc = hsum(west_plane(cur[rc + wl], c), c, east_plane(c, cur[rc + wr]));
            const HSum hb = hsum(west_plane(cur[rd + wl], b), b, east_plane(b, cur[rd + wr]));
            const uint32_t nxt = life_rule(ha, hc, hb, c);
            world_out[e * WPE + i] = nxt;
            pop += __builtin_popcount(nxt);
            mix_nibbles(nxt & ~c, nxt & c, mix[2 * i], mix[2 * i + 1]);
        }
        uint32_t *sp = reinterpret_cast<uint32_t *>(stable + e * SIZE);
        int32_t acc = 0;
        for (int q = 0; q < NCHUNK; ++q) {
            uint32_t s[4] = {sp[4 * q], sp[4 * q + 1], sp[4 * q + 2], sp[4 * q + 3]};
            if (act >= 0 && q == (act >> 4)) {
                const uint32_t m = 0xffu << ((act & 3) * 8);
                const int k = (act >> 2) & 3;
                s[k] = (s[k] & ~m) | (spawn4 & m);
            }
            const uint32_t m = mix[q];
            for (int k = 0; k < 4; ++k) {
                const uint32_t sv = (m >> (8 * k)) & 0xfu, bn = (m >> (8 * k + 4)) & 0xfu;
                s[k] = stable_update4(s[k], table[sv][0], table[bn][1], max4);
                for (int b = 0; b < 4; ++b) acc += (int8_t)(s[k] >> (8 * b));
                sp[4 * q + k] = s[k];
            }
        }
        if (reward_out) reward_out[e] = acc;
        if (alive_out) alive_out[e] = pop;
    }
    return err;
}

// mirrors stable_generic_kernel with the CGL_action+ dead-cell rules
void twin_stable_generic_rule(const uint32_t *prev, const uint32_t *next, int8_t *stable, uint64_t n_envs,
                              uint32_t side, int spawn, int stable_max, int rule, int empty, int empty_min)
{
    const uint32_t W = (side + 31) / 32;
    const uint64_t size = (uint64_t)side * side;
    for (uint64_t e = 0; e < n_envs; ++e)
        for (uint32_t r = 0; r < side; ++r)
            for (uint32_t c = 0; c < side; ++c) {
                const uint64_t widx = (e * side + r) * W + (c >> 5);
                const bool p = (prev[widx] >> (c & 31)) & 1u, n = (next[widx] >> (c & 31)) & 1u;
                int8_t &s = stable[e * size + (uint64_t)r * side + c];
                s = stable_update1_rule(rule, s, p, n, (int8_t)spawn, (int8_t)stable_max, (int8_t)empty,
                                        (int8_t)empty_min);
            }
}

// Exhaustive check of the 4-cells-per-word rule against the scalar rule: every stability byte in every
// lane, every transition (survive / born / dead), neighbours in the word holding other values.
uint64_t twin_rule_mismatches(int rule, int spawn, int stable_max, int empty, int empty_min)
{
    uint64_t bad = 0;
    const uint32_t spawn4 = rep4(spawn), max4 = rep4(stable_max), min4 = rep4(empty_min), empty4 = rep4(empty);
    for (int lane = 0; lane < 4; ++lane)
        for (int v = 0; v < 256; ++v)
            for (int tr = 0; tr < 3; ++tr)
                for (int other = 0; other < 3; ++other) {
                    uint32_t s = 0, surv = 0, born = 0;
                    int8_t want[4];
                    for (int b = 0; b < 4; ++b) {
                        const int8_t sv = (int8_t)(b == lane ? v : (v * 7 + 31 * b + 13) & 0xff);
                        const int t = b == lane ? tr : (other + b) % 3;
                        s |= (uint32_t)(uint8_t)sv << (8 * b);
                        if (t == 0) surv |= 0xffu << (8 * b);
                        if (t == 1) born |= 0xffu << (8 * b);
                        want[b] = stable_update1_rule(rule, sv, t == 0, t != 2, (int8_t)spawn, (int8_t)stable_max,
                                                      (int8_t)empty, (int8_t)empty_min);
                    }
                    const uint32_t got = stable_update4_rule(rule, s, surv, born, spawn4, max4, min4, empty4);
                    for (int b = 0; b < 4; ++b) bad += (int8_t)(got >> (8 * b)) != want[b];
                }
    return bad;
}

// Bit-sliced stability: transpose round trip and the sliced rule against the scalar rule, on `n_words`
// groups of 32 cells with the given int8 values and transitions (tr: 0 survive, 1 born, 2 dead).
uint64_t twin_sliced_mismatches(const int8_t *cells, const uint8_t *tr, uint64_t n_groups, int spawn, int stable_max)
{
    uint64_t bad = 0;
    for (uint64_t g = 0; g < n_groups; ++g) {
        uint32_t w[8], p[8], back[8], surv = 0, born = 0;
        memcpy(w, cells + 32 * g, 32);
        bytes_to_planes32(w, p);
        for (int j = 0; j < 32; ++j)
            for (int b = 0; b < 8; ++b)
                bad += ((p[b] >> j) & 1u) != (((uint8_t)cells[32 * g + j] >> b) & 1u);
        planes_to_bytes32(p, back);
        bad += memcmp(back, w, 32) != 0;
        for (int j = 0; j < 32; ++j) {
            if (tr[32 * g + j] == 0) surv |= 1u << j;
            if (tr[32 * g + j] == 1) born |= 1u << j;
        }
        // as the kernel does: spawn-relative planes, the relative rule, back to absolute values masked by "alive"
        add_const_sliced(p, -spawn);
        stable_update_sliced_rel(p, surv, (stable_max - spawn) & 0xff);
        add_const_sliced(p, spawn);
        for (int b = 0; b < 8; ++b) p[b] &= surv | born;
        planes_to_bytes32(p, back);
        for (int j = 0; j < 32; ++j) {
            const int8_t want = stable_update1(cells[32 * g + j], tr[32 * g + j] == 0, tr[32 * g + j] != 2,
                                               (int8_t)spawn, (int8_t)stable_max);
            bad += ((const int8_t *)back)[j] != want;
        }
    }
    return bad;
}

// The fork's decay rule on spawn-relative bit planes against the scalar rule (tr: 0 survive, 1 born, 2 dead).
uint64_t twin_sliced_decay_mismatches(const int8_t *cells, const uint8_t *tr, uint64_t n_groups, int spawn,
                                      int stable_max, int empty_min)
{
    uint64_t bad = 0;
    for (uint64_t g = 0; g < n_groups; ++g) {
        uint32_t w[8], p[8], back[8], surv = 0, born = 0;
        memcpy(w, cells + 32 * g, 32);
        bytes_to_planes32(w, p);
        for (int j = 0; j < 32; ++j) {
            if (tr[32 * g + j] == 0) surv |= 1u << j;
            if (tr[32 * g + j] == 1) born |= 1u << j;
        }
        add_const_sliced(p, -spawn);
        stable_update_sliced_decay(p, surv, born, (stable_max - spawn) & 0xff, (empty_min - spawn) & 0xff);
        add_const_sliced(p, spawn);
        planes_to_bytes32(p, back);
        for (int j = 0; j < 32; ++j) {
            const int8_t want = stable_update1_rule(CGL_DEAD_DECAY, cells[32 * g + j], tr[32 * g + j] == 0,
                                                    tr[32 * g + j] != 2, (int8_t)spawn, (int8_t)stable_max, 0,
                                                    (int8_t)empty_min);
            bad += ((const int8_t *)back)[j] != want;
        }
    }
    return bad;
}

// The fork's saturating rule on absolute bit planes against the scalar rule (tr: 0 survive, 1 born, 2 dead).
uint64_t twin_sliced_sat_mismatches(const int8_t *cells, const uint8_t *tr, uint64_t n_groups, int spawn, int stable_max,
                                    int empty, int empty_min)
{
    uint64_t bad = 0;
    for (uint64_t g = 0; g < n_groups; ++g) {
        uint32_t w[8], p[8], back[8], surv = 0, born = 0;
        memcpy(w, cells + 32 * g, 32);
        bytes_to_planes32(w, p);
        for (int j = 0; j < 32; ++j) {
            if (tr[32 * g + j] == 0) surv |= 1u << j;
            if (tr[32 * g + j] == 1) born |= 1u << j;
        }
        stable_update_sliced_sat(p, surv, born, spawn, stable_max, empty, empty_min);
        planes_to_bytes32(p, back);
        for (int j = 0; j < 32; ++j) {
            const int8_t want = stable_update1_rule(CGL_DEAD_SAT, cells[32 * g + j], tr[32 * g + j] == 0,
                                                    tr[32 * g + j] != 2, (int8_t)spawn, (int8_t)stable_max, (int8_t)empty,
                                                    (int8_t)empty_min);
            bad += ((const int8_t *)back)[j] != want;
        }
    }
    return bad;
}

// mirrors env_run_sliced_kernel<S, DECAY> (csrc/cgl_env_run.cu) for ONE env, side % 32 == 0: rows of the world
// and spawn-relative bit planes per row owner, horizontal sums (s0, s1) published once per row and step, the
// vote on the PREVIOUS step's change taken at the same point as the kernel's barrier.  Returns the steps executed.
int twin_env_run_sliced(uint32_t *world, int8_t *stable, uint32_t side, uint32_t max_steps, int stop_when_fixed,
                        int spawn, int stable_max, int decay, int empty_min)
{
    if (side % 32) return -1;
    const int S = (int)side, W = S / 32;
    std::vector<uint32_t> cw(world, world + S * W), pl((size_t)S * W * 8), h0(S * W), h1(S * W);
    const uint32_t *sb = reinterpret_cast<const uint32_t *>(stable);
    for (int r = 0; r < S; ++r)
        for (int w = 0; w < W; ++w) {
            uint32_t by[8], p[8];
            memcpy(by, sb + ((size_t)r * S + w * 32) / 4, 32);
            bytes_to_planes32(by, p);
            add_const_sliced(p, -spawn);
            memcpy(&pl[((size_t)r * W + w) * 8], p, 32);
        }
    const int max_rel = (stable_max - spawn) & 0xff, min_rel = (empty_min - spawn) & 0xff;
    uint32_t steps = 0, changed = 1;
    while (steps < max_steps) {
        std::vector<HSum> hs(S * W);
        for (int r = 0; r < S; ++r)
            for (int w = 0; w < W; ++w) {
                const uint32_t c = cw[r * W + w];
                hs[r * W + w] = hsum(west_plane(cw[r * W + (w + W - 1) % W], c), c, east_plane(c, cw[r * W + (w + 1) % W]));
                h0[r * W + w] = hs[r * W + w].s0;
                h1[r * W + w] = hs[r * W + w].s1;
            }
        if (stop_when_fixed && !changed) break;                 // the vote rides on the barrier that publishes the sums
        changed = 0;
        std::vector<uint32_t> nw(S * W);
        for (int r = 0; r < S; ++r)
            for (int w = 0; w < W; ++w) {
                const int ru = (r == 0 ? S - 1 : r - 1), rd = (r == S - 1 ? 0 : r + 1);
                const HSum up = {h0[ru * W + w], h1[ru * W + w], 0, 0}, dn = {h0[rd * W + w], h1[rd * W + w], 0, 0};
                const uint32_t c = cw[r * W + w];
                const uint32_t n = life_rule(up, hs[r * W + w], dn, c);
                changed |= n ^ c;
                uint32_t p[8];
                memcpy(p, &pl[((size_t)r * W + w) * 8], 32);
                if (decay) stable_update_sliced_decay(p, n & c, n & ~c, max_rel, min_rel);
                else stable_update_sliced_rel(p, n & c, max_rel);
                memcpy(&pl[((size_t)r * W + w) * 8], p, 32);
                nw[r * W + w] = n;
            }
        cw = nw;
        ++steps;
    }
    uint32_t *so = reinterpret_cast<uint32_t *>(stable);
    for (int r = 0; r < S; ++r)
        for (int w = 0; w < W; ++w) {
            uint32_t p[8], by[8];
            memcpy(p, &pl[((size_t)r * W + w) * 8], 32);
            add_const_sliced(p, spawn);
            if (!decay)
                for (int b = 0; b < 8; ++b) p[b] &= cw[r * W + w];
            planes_to_bytes32(p, by);
            memcpy(so + ((size_t)r * S + w * 32) / 4, by, 32);
            world[r * W + w] = cw[r * W + w];
        }
    return (int)steps;
}

// mirrors env_step_fused_kernel<S, IO, RULE >= 0>: the CGL_action+ variants (dead-cell rule, masked toggle)
int twin_env_step_fused_rule(const uint32_t *world_in, uint32_t *world_out, int8_t *stable, uint64_t n_envs,
                             uint32_t side, const int32_t *actions, int spawn, int stable_max, int rule, int empty,
                             int empty_min, int masked, int32_t *reward_out, uint32_t *alive_out)
{
    if (side % 32) return -1;
    const int S = (int)side, W = S / 32, WPE = S * W, SIZE = S * S, NCHUNK = SIZE / 16;
    const uint32_t spawn4 = rep4(spawn), max4 = rep4(stable_max), min4 = rep4(empty_min), empty4 = rep4(empty);
    uint32_t table[16][2];                       // [nibble][0] = byte mask, [1] = mask & SPAWN
    for (uint32_t i = 0; i < 16; ++i) { table[i][0] = nibble_to_bytemask(i); table[i][1] = table[i][0] & spawn4; }
    std::vector<uint32_t> cur(WPE), mix(2 * WPE);
    int err = 0;
    for (uint64_t e = 0; e < n_envs; ++e) {
        int act = -1;
        if (actions) {
            int a = actions[e];
            if (a >= 0 && a < SIZE) act = a;
            else if (a != SIZE) err = 1;
        }
        memcpy(cur.data(), world_in + e * WPE, WPE * 4);
        if (act >= 0) cur[act >> 5] ^= 1u << (act & 31);
        uint32_t pop = 0;
        for (int i = 0; i < WPE; ++i) {
            const int r = i / W, w = i % W;
            const int ru = (r == 0 ? S - 1 : r - 1) * W, rc = r * W, rd = (r == S - 1 ? 0 : r + 1) * W;
            const int wl = (w == 0 ? W - 1 : w - 1), wr = (w == W - 1 ? 0 : w + 1);
            const uint32_t a = cur[ru + w], c = cur[rc + w], b = cur[rd + w];
            const HSum ha = hsum(west_plane(cur[ru + wl], a), a, east_plane(a, cur[ru + wr]));
            const HSum hc = hsum(west_plane(cur[rc + wl], c), c, east_plane(c, cur[rc + wr]));
            const HSum hb = hsum(west_plane(cur[rd + wl], b), b, east_plane(b, cur[rd + wr]));
            const uint32_t nxt = life_rule(ha, hc, hb, c);
            world_out[e * WPE + i] = nxt;
            pop += __builtin_popcount(nxt);
            mix_nibbles(nxt & ~c, nxt & c, mix[2 * i], mix[2 * i + 1]);
        }
        uint32_t *sp = reinterpret_cast<uint32_t *>(stable + e * SIZE);
        int32_t acc = 0;
        for (int q = 0; q < NCHUNK; ++q) {
            uint32_t s[4] = {sp[4 * q], sp[4 * q + 1], sp[4 * q + 2], sp[4 * q + 3]};
            if (act >= 0 && q == (act >> 4)) {
                const uint32_t m = 0xffu << ((act & 3) * 8);
                const int k = (act >> 2) & 3;
                const bool alive_now = (cur[act >> 5] >> (act & 31)) & 1u;     // cur holds the toggled plane
                s[k] = (s[k] & ~m) | ((masked && !alive_now) ? 0u : (spawn4 & m));
            }
            const uint32_t m = mix[q];
            for (int k = 0; k < 4; ++k) {
                const uint32_t sv = (m >> (8 * k)) & 0xfu, bn = (m >> (8 * k + 4)) & 0xfu;
                s[k] = stable_update4_rule(rule, s[k], table[sv][0], table[bn][0], spawn4, max4, min4, empty4);
                for (int b = 0; b < 4; ++b) acc += (int8_t)(s[k] >> (8 * b));
                sp[4 * q + k] = s[k];
            }
        }
        if (reward_out) reward_out[e] = acc;
        if (alive_out) alive_out[e] = pop;
    }
    return err;
}

// the single-env server's packed-byte generation (cgl_bits.cuh life_next4_bytes), vectorised over n cases
void twin_life_next4_bytes(const uint32_t *u, const uint32_t *m, const uint32_t *d, const uint32_t *lc,
                           const uint32_t *rc, uint32_t *out, uint64_t n)
{
    for (uint64_t i = 0; i < n; ++i) out[i] = cgl::life_next4_bytes(u[i], m[i], d[i], lc[i], rc[i]);
}

}  // extern "C"
