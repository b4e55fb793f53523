/* TEST INFRASTRUCTURE ONLY -- see dpq_oracle.h.  Plain C restatement of the reference
 * algorithms; compiled with -ffp-contract=off so that no FMA is formed (the reference is
 * built for baseline x86-64, CMakeLists.txt:10). */
#include "dpq_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------ ADC table ------- */
/* DCAT.h:3750-3758: float accumulator, each term pow(float diff, 2) evaluated in double. */
void dpqo_lut(const float* cw, int M, int K, int Ds, const float* query, float* lut) {
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k) {
            float acc = 0.0f;
            const float* c = cw + ((size_t)m * K + k) * Ds;
            for (int d = 0; d < Ds; ++d) {
                float diff = c[d] - query[m * Ds + d];
                acc = (float)((double)acc + (double)diff * (double)diff);
            }
            lut[m * K + k] = acc;
        }
}

/* ------------------------------------------------------------------ top-k heap ------ */
/* Max-heap on float distance (DCAT.h:2054 cmp_max, :3853-3858 update rule). */
typedef struct {
    float d;
    uint32_t id;
} hent;

static void heap_sift_up(hent* h, int i) {
    while (i > 0) {
        int p = (i - 1) / 2;
        if (h[p].d < h[i].d) {
            hent t = h[p];
            h[p] = h[i];
            h[i] = t;
            i = p;
        } else
            break;
    }
}
static void heap_sift_down(hent* h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, b = i;
        if (l < n && h[b].d < h[l].d) b = l;
        if (r < n && h[b].d < h[r].d) b = r;
        if (b == i) break;
        hent t = h[b];
        h[b] = h[i];
        h[i] = t;
        i = b;
    }
}
/* "if size < k push; else if dist < top.first replace" with dist a double, top a float */
static void heap_offer(hent* h, int* n, int k, double dist, uint32_t id) {
    if (*n < k) {
        h[*n].d = (float)dist;
        h[*n].id = id;
        heap_sift_up(h, (*n)++);
    } else if (dist < (double)h[0].d) {
        h[0].d = (float)dist;
        h[0].id = id;
        heap_sift_down(h, *n, 0);
    }
}
/* pop into out[k-1..0] => ascending (DCAT.h:3884-3889). */
static void heap_drain(hent* h, int n, int k, int32_t* out_pos, float* out_dist) {
    for (int i = k - 1; i >= 0; --i) {
        if (n == 0) { /* fewer than k nodes: reference would pop an empty heap (UB) */
            out_pos[i] = -1;
            out_dist[i] = FLT_MAX;
            continue;
        }
        out_pos[i] = (int32_t)h[0].id;
        out_dist[i] = h[0].d;
        h[0] = h[--n];
        heap_sift_down(h, n, 0);
    }
}

/* ------------------------------------------------------------------ scan ------------ */
static inline int bitmap_bytes(int M) { return (M + 7) / 8; }

/* One node record: depth already known; reads bitmap + "to" bytes, updates the code stack
 * and the double distance exactly as DCAT.h:3797-3822. */
static inline double scan_node(const uint8_t* p, int64_t* off, int M, int K, const float* lut,
                               uint8_t* stack, double* dstack, int depth) {
    uint8_t* cur = stack + (size_t)depth * M;
    const uint8_t* par = stack + (size_t)(depth - 1) * M;
    memcpy(cur, par, (size_t)M);
    double dist = dstack[depth - 1];
    uint32_t bitmap = 0;
    for (int b = 0; b < bitmap_bytes(M); ++b) bitmap |= (uint32_t)p[(*off)++] << (8 * b);
    for (int m = 0; m < M; ++m)
        if ((bitmap >> m) & 1) {
            uint8_t cid = p[(*off)++];
            cur[m] = cid;
            dist -= (double)lut[m * K + par[m]];
            dist += (double)lut[m * K + cid];
        }
    dstack[depth] = dist;
    return dist;
}

int64_t dpqo_scan(const uint8_t* payload, int64_t n_bytes, int64_t n_codes, int M, int K,
                  const float* lut, int topk, int32_t* out_pos, float* out_dist,
                  float* node_dist) {
    (void)n_bytes;
    int maxd = M > 8 ? 16 : 8; /* depth nibble range */
    uint8_t* stack = (uint8_t*)calloc((size_t)(maxd + 1) * M, 1);
    double* dstack = (double*)calloc((size_t)maxd + 1, sizeof(double));
    hent* heap = (hent*)malloc(sizeof(hent) * (size_t)(topk > 0 ? topk : 1));
    int hn = 0;
    int64_t off = 0;
    int dmask = M > 8 ? 15 : 7; /* reference masks with &7 (DCAT.h:3794,3827) */

    double qdist = 0; /* root: M raw bytes, double sum (DCAT.h:3773-3781) */
    for (int m = 0; m < M; ++m) {
        uint8_t cid = payload[off++];
        qdist += (double)lut[m * K + cid];
        stack[m] = cid;
    }
    dstack[0] = qdist;
    if (n_codes > 0) {
        heap_offer(heap, &hn, topk, qdist, 0);
        if (node_dist) node_dist[0] = (float)qdist;
    }
    int64_t i = 1;
    for (; i + 1 < n_codes; i += 2) {
        int depths = payload[off++];
        double d1 = scan_node(payload, &off, M, K, lut, stack, dstack, depths & dmask);
        heap_offer(heap, &hn, topk, d1, (uint32_t)i);
        double d2 = scan_node(payload, &off, M, K, lut, stack, dstack, (depths >> 4) & dmask);
        heap_offer(heap, &hn, topk, d2, (uint32_t)(i + 1));
        if (node_dist) {
            node_dist[i] = (float)d1;
            node_dist[i + 1] = (float)d2;
        }
    }
    if (i == n_codes - 1) { /* trailing node: depth byte unmasked (DCAT.h:3861) */
        int depth = payload[off++];
        double d = scan_node(payload, &off, M, K, lut, stack, dstack, depth);
        heap_offer(heap, &hn, topk, d, (uint32_t)i); /* reference pushes i+1: App. C.1 */
        if (node_dist) node_dist[i] = (float)d;
    }
    heap_drain(heap, hn, topk, out_pos, out_dist);
    free(stack);
    free(dstack);
    free(heap);
    return off;
}

int64_t dpqo_decode(const uint8_t* payload, int64_t n_bytes, int64_t n_codes, int M,
                    uint8_t* codes, uint8_t* depth, int32_t* parent) {
    (void)n_bytes;
    int maxd = M > 8 ? 16 : 8;
    int dmask = M > 8 ? 15 : 7;
    uint8_t* stack = (uint8_t*)calloc((size_t)(maxd + 1) * M, 1);
    int32_t* pstack = (int32_t*)calloc((size_t)maxd + 1, sizeof(int32_t));
    int64_t off = 0;
    for (int m = 0; m < M; ++m) stack[m] = payload[off++];
    if (n_codes > 0) {
        if (codes) memcpy(codes, stack, (size_t)M);
        if (depth) depth[0] = 0;
        if (parent) parent[0] = -1;
    }
    pstack[0] = 0;
    int depths = 0;
    for (int64_t i = 1; i < n_codes; ++i) {
        int d;
        if (i & 1) {
            depths = payload[off++];
            d = (i == n_codes - 1) ? depths : (depths & dmask);
        } else
            d = (depths >> 4) & dmask;
        uint8_t* cur = stack + (size_t)d * M;
        memcpy(cur, stack + (size_t)(d - 1) * M, (size_t)M);
        uint32_t bitmap = 0;
        for (int b = 0; b < bitmap_bytes(M); ++b) bitmap |= (uint32_t)payload[off++] << (8 * b);
        for (int m = 0; m < M; ++m)
            if ((bitmap >> m) & 1) cur[m] = payload[off++];
        if (codes) memcpy(codes + (size_t)i * M, cur, (size_t)M);
        if (depth) depth[i] = (uint8_t)d;
        if (parent) parent[i] = pstack[d - 1];
        pstack[d] = (int32_t)i;
    }
    free(stack);
    free(pstack);
    return off;
}

/* ------------------------------------------------------------------ encode ---------- */
/* pq_tree.cpp:215-237: sequential float, separate multiply and add, strict <. */
void dpqo_encode(const float* cw, int M, int K, int Ds, const float* x, int64_t n, int D,
                 uint8_t* codes) {
    float* v = (float*)malloc(sizeof(float) * (size_t)M * Ds);
    for (int64_t i = 0; i < n; ++i) {
        for (int d = 0; d < M * Ds; ++d) v[d] = d < D ? x[(size_t)i * D + d] : 0.0f;
        for (int m = 0; m < M; ++m) {
            float min_dist = FLT_MAX;
            int min_ks = -1;
            for (int ks = 0; ks < K; ++ks) {
                float dist = 0;
                const float* c = cw + ((size_t)m * K + ks) * Ds;
                for (int ds = 0; ds < Ds; ++ds) {
                    float diff = v[m * Ds + ds] - c[ds];
                    dist += diff * diff;
                }
                if (dist < min_dist) {
                    min_dist = dist;
                    min_ks = ks;
                }
            }
            codes[(size_t)i * M + m] = (uint8_t)min_ks;
        }
    }
    free(v);
}

/* ------------------------------------------------------------------ edge search ----- */
typedef struct {
    u128 key;
    uint32_t pos;
} kent;

/* Stable LSD radix sort of (key, pos) by key; skips bytes on which all keys agree. */
static void radix_sort_keys(kent* a, kent* tmp, int64_t n) {
    for (int byte = 0; byte < 16; ++byte) {
        int64_t cnt[257];
        memset(cnt, 0, sizeof cnt);
        int sh = byte * 8;
        for (int64_t i = 0; i < n; ++i) cnt[((unsigned)(a[i].key >> sh) & 255u) + 1]++;
        int skip = 0;
        for (int b = 0; b < 256; ++b)
            if (cnt[b + 1] == n) skip = 1;
        if (skip) continue;
        for (int b = 0; b < 256; ++b) cnt[b + 1] += cnt[b];
        for (int64_t i = 0; i < n; ++i) tmp[cnt[(unsigned)(a[i].key >> sh) & 255u]++] = a[i];
        memcpy(a, tmp, sizeof(kent) * (size_t)n);
    }
}

typedef struct {
    uint32_t* v;
    int64_t n, cap;
} uvec;
static void uvec_push(uvec* u, uint32_t x) {
    if (u->n == u->cap) {
        u->cap = u->cap ? u->cap * 2 : 1024;
        u->v = (uint32_t*)realloc(u->v, sizeof(uint32_t) * (size_t)u->cap);
    }
    u->v[u->n++] = x;
}

/* next selector in std::prev_permutation order over a bool vector (CT.h:75-90). */
static int prev_perm(uint8_t* s, int n) {
    int i = n - 1;
    while (i > 0 && s[i - 1] <= s[i]) --i;
    if (i <= 0) return 0;
    int j = n - 1;
    while (s[j] >= s[i - 1]) --j;
    uint8_t t = s[i - 1];
    s[i - 1] = s[j];
    s[j] = t;
    for (int a = i, b = n - 1; a < b; ++a, --b) {
        t = s[a];
        s[a] = s[b];
        s[b] = t;
    }
    return 1;
}

int64_t dpqo_find_edges(const uint8_t* codes, int64_t n_codes, int M, int K,
                        int max_height_folds, int method, uint32_t* edges, uint32_t* root_id) {
    int LOG_K = (int)round(log2((double)K)); /* DCAT.h:454 */
    int MAX_HEIGHT = M * max_height_folds;   /* DCAT.h:1262 */
    int64_t n_edges = 0;
    uint8_t* heights = (uint8_t*)calloc((size_t)n_codes, 1);
    uint8_t* is_active = (uint8_t*)malloc((size_t)n_codes);
    memset(is_active, 1, (size_t)n_codes);
    uint32_t* ids = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n_codes);
    int64_t n_ids = n_codes;
    for (int64_t i = 0; i < n_codes; ++i) ids[i] = (uint32_t)i;
    uvec finalists = {0, 0, 0};
    kent* ha = (kent*)malloc(sizeof(kent) * (size_t)(n_codes ? n_codes : 1));
    kent* tmp = (kent*)malloc(sizeof(kent) * (size_t)(n_codes ? n_codes : 1));
    uint8_t* is_merged = (uint8_t*)malloc((size_t)(n_codes ? n_codes : 1));
    uint8_t sel[64];
    *root_id = 0;

    for (int diff = 0; diff <= M; ++diff) { /* dmain:126 forces diff_argument = M */
        memset(is_merged, 0, (size_t)n_ids);
        for (int i = 0; i < M; ++i) sel[i] = i < M - diff;
        do { /* one kept-subspace combination (DCAT.h:481-600) */
            int64_t na = 0;
            for (int64_t l = 0; l < n_ids; ++l) {
                if (is_merged[l]) continue;
                uint32_t code_id = ids[l];
                u128 key = 0;
                for (int m = 0; m < M; ++m)
                    if (sel[m]) key |= (u128)codes[(size_t)code_id * M + m] << (LOG_K * m);
                ha[na].key = key;
                ha[na].pos = (uint32_t)l;
                ++na;
            }
            radix_sort_keys(ha, tmp, na);
            for (int64_t i = 0; i < na; ++i) {
                int64_t end = i + 1;
                while (end < na && ha[end].key == ha[i].key) ++end;
                if (end == i + 1) continue;
                uint32_t parent_pos, parent_code;
                if (method == 2) { /* DCAT.h:731-743 */
                    parent_pos = ha[i].pos;
                    parent_code = ids[parent_pos];
                    for (int64_t j = i + 1; j < end; ++j) {
                        uint32_t c = ids[ha[j].pos];
                        if ((int)heights[c] + 1 > (int)heights[parent_code])
                            heights[parent_code] = (uint8_t)(heights[c] + 1);
                    }
                    if ((int)heights[parent_code] >= MAX_HEIGHT - 2) {
                        uvec_push(&finalists, parent_code);
                        is_merged[parent_pos] = 1;
                    }
                } else { /* DCAT.h:545-575 */
                    int max_height = -1;
                    parent_pos = 0;
                    for (int64_t j = i; j < end; ++j) {
                        uint32_t c = ids[ha[j].pos];
                        if ((int)heights[c] > max_height) {
                            max_height = heights[c];
                            parent_pos = ha[j].pos;
                        }
                    }
                    parent_code = ids[parent_pos];
                    int second = 0;
                    for (int64_t j = i; j < end; ++j) {
                        uint32_t c = ids[ha[j].pos];
                        if (c == parent_code) continue;
                        if ((int)heights[c] > second) second = heights[c];
                    }
                    if (second == max_height) heights[parent_code]++;
                    max_height++;
                    if (max_height >= MAX_HEIGHT - 2) {
                        uvec_push(&finalists, parent_code);
                        is_merged[parent_pos] = 1;
                    }
                }
                *root_id = parent_code;
                for (int64_t j = i; j < end; ++j) {
                    uint32_t pos = ha[j].pos;
                    if (pos == parent_pos) continue;
                    uint32_t c = ids[pos];
                    is_merged[pos] = 1;
                    if (method != 2) is_active[c] = 0;
                    edges[2 * n_edges] = parent_code;
                    edges[2 * n_edges + 1] = c;
                    ++n_edges;
                }
                i = end - 1;
            }
        } while (prev_perm(sel, M));
        int64_t nn = 0; /* next round's ids (DCAT.h:611-615, 1283-1286) */
        for (int64_t l = 0; l < n_ids; ++l)
            if (!is_merged[l]) ids[nn++] = ids[l];
        n_ids = nn;
        if (n_ids <= 1) break;
    }
    if (n_ids > 0) uvec_push(&finalists, ids[0]); /* DCAT.h:1292-1294 */
    if (finalists.n > 0) {                        /* DCAT.h:1297-1313 */
        uint32_t p = finalists.v[0];
        *root_id = p;
        for (int64_t i = 1; i < finalists.n; ++i) {
            edges[2 * n_edges] = p;
            edges[2 * n_edges + 1] = finalists.v[i];
            ++n_edges;
        }
    }
    free(heights);
    free(is_active);
    free(ids);
    free(finalists.v);
    free(ha);
    free(tmp);
    free(is_merged);
    return n_edges;
}

/* ------------------------------------------------------------------ layout ---------- */
void dpqo_centroid_tables(const float* cw, int M, int K, int Ds, float* tables) {
    for (int m = 0; m < M; ++m)
        for (int j = 0; j < K; ++j)
            for (int k = 0; k < K; ++k) {
                float dist = 0;
                for (int d = 0; d < Ds; ++d) {
                    float diff = cw[((size_t)m * K + j) * Ds + d] - cw[((size_t)m * K + k) * Ds + d];
                    dist = (float)((double)dist + (double)diff * (double)diff);
                }
                tables[((size_t)m * K + j) * K + k] = dist;
            }
}

/* CT.h:827-835 */
static float table_dist(const uint8_t* codes, int M, int K, const float* tables, uint32_t a,
                        uint32_t b) {
    float sum = 0;
    for (int m = 0; m < M; ++m) {
        int ca = codes[(size_t)a * M + m], cb = codes[(size_t)b * M + m];
        sum += tables[((size_t)m * K + ca) * K + cb];
    }
    return sum;
}

/* stable insertion/merge sort of ids by key descending */
static void stable_sort_desc(uint32_t* a, int64_t n, const float* key, uint32_t* tmp) {
    if (n < 2) return;
    if (n <= 16) {
        for (int64_t i = 1; i < n; ++i) {
            uint32_t x = a[i];
            int64_t j = i;
            while (j > 0 && key[x] > key[a[j - 1]]) {
                a[j] = a[j - 1];
                --j;
            }
            a[j] = x;
        }
        return;
    }
    int64_t h = n / 2;
    stable_sort_desc(a, h, key, tmp);
    stable_sort_desc(a + h, n - h, key, tmp);
    int64_t i = 0, j = h, k = 0;
    while (i < h && j < n) tmp[k++] = (key[a[j]] > key[a[i]]) ? a[j++] : a[i++];
    while (i < h) tmp[k++] = a[i++];
    while (j < n) tmp[k++] = a[j++];
    memcpy(a, tmp, sizeof(uint32_t) * (size_t)n);
}

void dpqo_layout(const uint8_t* codes, int64_t n_codes, int M, int K, const uint32_t* edges,
                 uint32_t root_id, const float* tables, uint32_t* vec_id, uint32_t* parent_pos,
                 uint32_t* child_num, uint8_t* depth, float* max_dist, float* max_dist2p) {
    int64_t N = n_codes, E = n_codes > 0 ? n_codes - 1 : 0;
    uint32_t* parents = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(N + 1));
    for (int64_t i = 0; i < N; ++i) parents[i] = 0xFFFFFFFFu;
    for (int64_t e = 0; e < E; ++e) parents[edges[2 * e + 1]] = edges[2 * e];
    /* CSR, children in emission order == stable sort by parent (DCAT.h:1077-1104) */
    uint32_t* offsets = (uint32_t*)calloc((size_t)N + 2, sizeof(uint32_t));
    uint32_t* row = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(E + 1));
    for (int64_t e = 0; e < E; ++e) offsets[edges[2 * e] + 1]++;
    for (int64_t i = 0; i < N; ++i) offsets[i + 1] += offsets[i];
    uint32_t* fill = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(N + 1));
    memcpy(fill, offsets, sizeof(uint32_t) * (size_t)(N + 1));
    for (int64_t e = 0; e < E; ++e) row[fill[edges[2 * e]]++] = edges[2 * e + 1];
    /* max distance to ancestors, at most 16 levels (DCAT.h:1396-1417) */
    float* md = (float*)calloc((size_t)N + 1, sizeof(float));
    float* md2p = (float*)calloc((size_t)N + 1, sizeof(float));
    for (int64_t vid = 0; vid < N; ++vid) {
        uint32_t parent = parents[vid], prev = (uint32_t)vid;
        int d = 0;
        while (parent != 0xFFFFFFFFu) {
            if (d++ >= 16) break;
            float dist = table_dist(codes, M, K, tables, (uint32_t)vid, parent);
            if (dist > md[parent]) md[parent] = dist;
            if (dist > md2p[prev]) md2p[prev] = dist;
            prev = parent;
            parent = parents[parent];
        }
    }
    /* children by max_dist2p descending, stable (DCAT.h:1421-1426) */
    for (int64_t vid = 0; vid < N; ++vid)
        stable_sort_desc(row + offsets[vid], (int64_t)offsets[vid + 1] - offsets[vid], md2p, fill);
    /* pre-order DFS (DCAT.h:1156-1183), iterative */
    uint32_t* st_vid = (uint32_t*)malloc(sizeof(uint32_t) * 64);
    uint32_t* st_it = (uint32_t*)malloc(sizeof(uint32_t) * 64);
    uint32_t* st_pos = (uint32_t*)malloc(sizeof(uint32_t) * 64);
    int64_t cap = 64;
    uint32_t* o_vid = vec_id ? vec_id : (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(N + 1));
    uint32_t* o_cn = child_num ? child_num : (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(N + 1));
    if (N > 0) {
        int64_t sp = 0;
        uint32_t node_id = 0;
        o_vid[0] = root_id;
        if (parent_pos) parent_pos[0] = 0xFFFFFFFFu;
        if (depth) depth[0] = 0;
        st_vid[0] = root_id;
        st_it[0] = offsets[root_id];
        st_pos[0] = 0;
        while (sp >= 0) {
            uint32_t pv = st_vid[sp];
            if (st_it[sp] < offsets[pv + 1]) {
                uint32_t child = row[st_it[sp]++];
                ++node_id;
                o_vid[node_id] = child;
                if (parent_pos) parent_pos[node_id] = st_pos[sp];
                if (depth) depth[node_id] = (uint8_t)(sp + 1);
                if (sp + 2 > cap) {
                    cap *= 2;
                    st_vid = (uint32_t*)realloc(st_vid, sizeof(uint32_t) * (size_t)cap);
                    st_it = (uint32_t*)realloc(st_it, sizeof(uint32_t) * (size_t)cap);
                    st_pos = (uint32_t*)realloc(st_pos, sizeof(uint32_t) * (size_t)cap);
                }
                ++sp;
                st_vid[sp] = child;
                st_it[sp] = offsets[child];
                st_pos[sp] = node_id;
            } else {
                o_cn[st_pos[sp]] = node_id - st_pos[sp];
                --sp;
            }
        }
    }
    for (int64_t pos = 0; pos < N; ++pos) {
        if (max_dist) max_dist[pos] = sqrtf(md[o_vid[pos]]);
        if (max_dist2p) max_dist2p[pos] = sqrtf(md2p[o_vid[pos]]);
    }
    if (!vec_id) free(o_vid);
    if (!child_num) free(o_cn);
    free(parents);
    free(offsets);
    free(row);
    free(fill);
    free(md);
    free(md2p);
    free(st_vid);
    free(st_it);
    free(st_pos);
}

void dpqo_qnodes8(const uint8_t* codes, int64_t n_codes, const uint32_t* vec_id,
                  const uint32_t* parent_pos, const uint32_t* child_num, const uint8_t* depth,
                  const float* max_dist, const float* max_dist2p, dpqo_qnode8* nodes) {
    const int M = 8;
    memset(nodes, 0, sizeof(dpqo_qnode8) * (size_t)(n_codes + 1));
    for (int64_t i = 0; i <= n_codes; ++i) nodes[i].sub_tree_size = 1;
    for (int64_t pos = 0; pos < n_codes; ++pos) {
        dpqo_qnode8* q = nodes + pos;
        q->vec_id = vec_id[pos];
        q->parent_pos = parent_pos[pos];
        q->child_pos_start = (uint32_t)pos + 1;
        q->child_num = child_num[pos];
        q->max_dist = max_dist[pos];
        q->max_dist2p = max_dist2p[pos];
        q->depth = depth[pos];
        const uint8_t* c = codes + (size_t)vec_id[pos] * M;
        if (pos == 0) { /* DCAT.h:1437-1445 */
            for (int m = 0; m < M; ++m) {
                q->diffs[m][0] = (uint8_t)m;
                q->diffs[m][1] = 0xFF;
                q->diffs[m][2] = c[m];
            }
            q->diff_num = (uint8_t)M;
        } else {
            const uint8_t* p = codes + (size_t)vec_id[parent_pos[pos]] * M;
            int nd = 0;
            for (int m = 0; m < M; ++m)
                if (p[m] != c[m]) {
                    q->diffs[nd][0] = (uint8_t)m;
                    q->diffs[nd][1] = p[m];
                    q->diffs[nd][2] = c[m];
                    ++nd;
                }
            q->diff_num = (uint8_t)nd;
        }
    }
}

/* ------------------------------------------------------------------ stream ---------- */
int64_t dpqo_stream_bytes(const uint8_t* codes, int64_t n_codes, int M, const uint32_t* vec_id,
                          const uint32_t* parent_pos) {
    int64_t n_diffs = 0;
    for (int64_t pos = 1; pos < n_codes; ++pos) {
        const uint8_t* c = codes + (size_t)vec_id[pos] * M;
        const uint8_t* p = codes + (size_t)vec_id[parent_pos[pos]] * M;
        for (int m = 0; m < M; ++m) n_diffs += p[m] != c[m];
    }
    if (M == 8) return 8 + n_diffs + (3 * (n_codes - 1) + 1) / 2; /* DCAT.h:1765 */
    return M + n_diffs + (int64_t)bitmap_bytes(M) * (n_codes - 1) + (n_codes - 1 + 1) / 2;
}

int64_t dpqo_stream(const uint8_t* codes, int64_t n_codes, int M, const uint32_t* vec_id,
                    const uint32_t* parent_pos, const uint8_t* depth, uint8_t* payload) {
    int64_t off = 0;
    if (n_codes <= 0) return 0;
    for (int m = 0; m < M; ++m) payload[off++] = codes[(size_t)vec_id[0] * M + m];
    for (int64_t i = 1; i < n_codes; ++i) {
        if (i & 1) {
            uint8_t d = depth[i];
            if (i + 1 < n_codes) d = (uint8_t)(d | (depth[i + 1] << 4));
            payload[off++] = d;
        }
        const uint8_t* c = codes + (size_t)vec_id[i] * M;
        const uint8_t* p = codes + (size_t)vec_id[parent_pos[i]] * M;
        uint32_t bitmap = 0;
        for (int m = 0; m < M; ++m)
            if (p[m] != c[m]) bitmap |= 1u << m;
        for (int b = 0; b < bitmap_bytes(M); ++b) payload[off++] = (uint8_t)(bitmap >> (8 * b));
        for (int m = 0; m < M; ++m)
            if (p[m] != c[m]) payload[off++] = c[m];
    }
    return off;
}

/* ------------------------------------------------------------------ ground truth ---- */
/* pmain:138-166: float difference, float product, double sum; heap holds floats. */
void dpqo_groundtruth_chunk(const float* base, int64_t n, int64_t id0, const float* queries,
                            int Q, int D, int topk, float* heap_dist, uint32_t* heap_id,
                            int32_t* count) {
    hent* h = (hent*)malloc(sizeof(hent) * (size_t)topk);
    for (int q = 0; q < Q; ++q) {
        int hn = count[q];
        for (int i = 0; i < hn; ++i) {
            h[i].d = heap_dist[(size_t)q * topk + i];
            h[i].id = heap_id[(size_t)q * topk + i];
        }
        const float* qu = queries + (size_t)q * D;
        for (int64_t it = 0; it < n; ++it) {
            const float* v = base + (size_t)it * D;
            double distance = 0.0;
            for (int d = 0; d < D; ++d) distance += (double)((v[d] - qu[d]) * (v[d] - qu[d]));
            heap_offer(h, &hn, topk, distance, (uint32_t)(id0 + it));
        }
        count[q] = hn;
        for (int i = 0; i < hn; ++i) {
            heap_dist[(size_t)q * topk + i] = h[i].d;
            heap_id[(size_t)q * topk + i] = h[i].id;
        }
    }
    free(h);
}

void dpqo_groundtruth_finish(int Q, int topk, float* heap_dist, uint32_t* heap_id,
                             const int32_t* count, uint32_t* out_id, float* out_dist) {
    hent* h = (hent*)malloc(sizeof(hent) * (size_t)topk);
    int32_t* pos = (int32_t*)malloc(sizeof(int32_t) * (size_t)topk);
    for (int q = 0; q < Q; ++q) {
        int hn = count[q];
        for (int i = 0; i < hn; ++i) {
            h[i].d = heap_dist[(size_t)q * topk + i];
            h[i].id = heap_id[(size_t)q * topk + i];
        }
        heap_drain(h, hn, topk, pos, out_dist + (size_t)q * topk);
        for (int i = 0; i < topk; ++i) out_id[(size_t)q * topk + i] = (uint32_t)pos[i];
    }
    free(h);
    free(pos);
}
