/*
 * dsim_synth.cpp — synthetic hex grids and family populations (see include/dsim_synth.h).
 * Host-only; shares nothing with oracle/.
 */
#include "../../include/dsim_synth.h"
#include "../csrc/dsim_rng.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

/* the 18 axial offsets at hex distance 1 and 2 */
struct Off { int dq, dr, ring; };
std::vector<Off> ring_offsets() {
    std::vector<Off> v;
    for (int dq = -2; dq <= 2; ++dq)
        for (int dr = -2; dr <= 2; ++dr) {
            const int d = (std::abs(dq) + std::abs(dr) + std::abs(dq + dr)) / 2;
            if (d == 1 || d == 2) v.push_back({dq, dr, d});
        }
    return v;
}

/* odd-r offset <-> axial */
inline int axial_q(int col, int row) { return col - (row - (row & 1)) / 2; }
inline int offset_col(int q, int row) { return q + (row - (row & 1)) / 2; }

struct Edge { uint32_t dst; double sec; };

void node_edges(uint32_t W, uint32_t H, dsim_stream s, const std::vector<Off> &offs, uint32_t node, std::vector<Edge> &out) {
    out.clear();
    const int row = (int)(node / W), col = (int)(node % W);
    const int q = axial_q(col, row);
    for (const Off &o : offs) {
        const int r2 = row + o.dr;
        if (r2 < 0 || r2 >= (int)H) continue;
        const int c2 = offset_col(q + o.dq, r2);
        if (c2 < 0 || c2 >= (int)W) continue;
        const uint32_t dst = (uint32_t)r2 * W + (uint32_t)c2;
        const dsim_u4 d = dsim_draw(s, node, dst, DSIM_P_SYN_EDGE, 0);
        out.push_back({dst, 14400.0 * o.ring * (1.0 + 0.25 * dsim_u53(d.x, d.y))});
    }
    std::sort(out.begin(), out.end(), [](const Edge &a, const Edge &b) { return a.dst < b.dst; });
}

} // namespace

extern "C" {

int dsim_synth_grid_sizes(uint32_t W, uint32_t H, uint64_t *n_nodes, uint64_t *n_edges) {
    if (W == 0 || H == 0 || (uint64_t)W * H > 0xFFFFFFF0ull) return DSIM_ERR_INVALID;
    const std::vector<Off> offs = ring_offsets();
    uint64_t e = 0;
    /* interior rows/cols have all 18; count exactly, per row parity and column */
    for (uint32_t row = 0; row < H; ++row) {
        for (uint32_t col = 0; col < W; ++col) {
            /* fast path for the interior */
            if (row >= 2 && row + 2 < H && col >= 2 && col + 2 < W) { e += 18; continue; }
            const int q = axial_q((int)col, (int)row);
            for (const Off &o : offs) {
                const int r2 = (int)row + o.dr;
                if (r2 < 0 || r2 >= (int)H) continue;
                const int c2 = offset_col(q + o.dq, r2);
                if (c2 < 0 || c2 >= (int)W) continue;
                ++e;
            }
        }
    }
    if (n_nodes) *n_nodes = (uint64_t)W * H;
    if (n_edges) *n_edges = e;
    return DSIM_OK;
}

int dsim_synth_grid(uint32_t W, uint32_t H, uint64_t seed, double recovery, uint64_t *h3, double *lon, double *lat,
                    uint64_t *eco_off, uint32_t *eco_id, double *popcap, uint64_t *edge_off, uint32_t *edge_dst,
                    double *edge_sec, double *res, double *max_res) {
    if (W == 0 || H == 0 || !eco_off || !eco_id || !edge_off || !edge_dst || !edge_sec) return DSIM_ERR_INVALID;
    const dsim_stream s = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const std::vector<Off> offs = ring_offsets();
    const uint64_t n = (uint64_t)W * H;
    std::vector<Edge> tmp;
    uint64_t e = 0;
    for (uint64_t v = 0; v < n; ++v) {
        const uint32_t row = (uint32_t)(v / W), col = (uint32_t)(v % W);
        if (h3) h3[v] = v;
        /* lon -168.6..-34.5, lat -56.0..74.5 as in src/serialize_graph.rs:146-149 */
        if (lon) lon[v] = -168.6 + (col + 0.5 * (row & 1)) * (134.1 / W);
        if (lat) lat[v] = -56.0 + row * (130.5 / H);
        eco_off[v] = v;
        eco_id[v] = (uint32_t)((uint64_t)row * 16 / H);
        const dsim_u4 d = dsim_draw(s, 0, v, DSIM_P_SYN_POPCAP, 0);
        const double u1 = 1.0 - dsim_u53(d.x, d.y), u2 = dsim_u53(d.z, d.w);
        const double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
        const double pc = std::min(400.0, std::max(0.5, 20.0 * std::exp(z)));
        if (popcap) popcap[v] = pc;
        if (max_res) max_res[v] = pc / recovery;
        if (res) res[v] = pc / recovery;
        edge_off[v] = e;
        node_edges(W, H, s, offs, (uint32_t)v, tmp);
        for (const Edge &x : tmp) {
            edge_dst[e] = x.dst;
            edge_sec[e] = x.sec;
            ++e;
        }
    }
    eco_off[n] = n;
    edge_off[n] = e;
    return DSIM_OK;
}

int dsim_synth_families(uint32_t W, uint32_t H, uint64_t seed, uint64_t n, uint64_t n_occupied, uint32_t hist_fill,
                        dsim_families *f) {
    if (W == 0 || H == 0 || !f) return DSIM_ERR_INVALID;
    const dsim_stream s = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const uint64_t P = (uint64_t)W * H;
    uint64_t seed_culture[64];
    for (uint32_t b = 0; b < 64; ++b) {
        const dsim_u4 d = dsim_draw(s, 0, b, DSIM_P_SYN_CULTURE, 0);
        seed_culture[b] = ((uint64_t)d.y << 32) | d.x;
    }
    f->n = n;
    for (uint64_t i = 0; i < n; ++i) {
        const dsim_u4 a = dsim_draw(s, 0, i, DSIM_P_SYN_FAMILY, 0);
        const dsim_u4 b = dsim_draw(s, 0, i, DSIM_P_SYN_FAMILY, 1);
        uint32_t loc;
        if (n_occupied == 0) {
            loc = (uint32_t)(((unsigned __int128)(((uint64_t)a.y << 32) | a.x) * P) >> 64);
        } else {
            const double u = dsim_u53(a.x, a.y);
            uint64_t k = (uint64_t)((double)n_occupied * std::pow(u, 1.25));
            if (k >= n_occupied) k = n_occupied - 1;
            const dsim_u4 site = dsim_draw(s, 0, k, DSIM_P_SYN_SITE, 0);
            loc = (uint32_t)(((unsigned __int128)(((uint64_t)site.y << 32) | site.x) * P) >> 64);
        }
        const uint32_t row = loc / W, col = loc % W;
        if (f->id) f->id[i] = i;
        if (f->parent_id) f->parent_id[i] = DSIM_NO_PARENT;
        if (f->birth_index) f->birth_index[i] = 0;
        if (f->location) f->location[i] = loc;
        if (f->size) f->size[i] = 2 + dsim_mulhi32(a.z, 8);
        if (f->stored) f->stored[i] = 5.0 * dsim_u53(b.x, b.y);
        if (f->last_harvest) f->last_harvest[i] = 0.0;
        if (f->till_adult) f->till_adult[i] = dsim_mulhi32(a.w, 18);
        if (f->till_mutation) f->till_mutation[i] = 0;
        if (f->has_mutation_clock) f->has_mutation_clock[i] = 0;
        if (f->n_offspring) f->n_offspring[i] = 0;
        if (f->culture) {
            const uint32_t block = (uint32_t)((uint64_t)row * 8 / H) * 8 + (uint32_t)((uint64_t)col * 8 / W);
            uint64_t c = seed_culture[block];
            const uint32_t flips = b.z >> 30; /* 0..3 */
            for (uint32_t k = 0; k < flips; ++k) c ^= (uint64_t)1 << ((b.w >> (6 * k)) & 63);
            f->culture[i] = c;
        }
        if (f->hist_off) {
            f->hist_off[i] = i * hist_fill;
            for (uint32_t k = 0; k < hist_fill; ++k) {
                const dsim_u4 hh = dsim_draw(s, k, i, DSIM_P_SYN_HIST, 0);
                int r2 = (int)row + (int)dsim_mulhi32(hh.x, 5) - 2;
                int c2 = (int)col + (int)dsim_mulhi32(hh.y, 5) - 2;
                r2 = std::min((int)H - 1, std::max(0, r2));
                c2 = std::min((int)W - 1, std::max(0, c2));
                if (f->hist_patch) f->hist_patch[i * hist_fill + k] = (uint32_t)r2 * W + (uint32_t)c2;
                if (f->hist_harvest) f->hist_harvest[i * hist_fill + k] = 0.4 * dsim_u53(hh.z, hh.w);
            }
        }
        if (f->adapt_off) f->adapt_off[i] = 0;
    }
    if (f->hist_off) f->hist_off[n] = n * hist_fill;
    if (f->adapt_off) f->adapt_off[n] = 0;
    return DSIM_OK;
}

} /* extern "C" */
