// S1 on sm_100a: threshold + 8-connected blobs + contour-polygon moments.
//
// Replaces Cameras._find_dot (reference computer_code/api/helpers.py:143-163):
//   grey = cvtColor(img, RGB2GRAY); binary = grey > 51; contours = findContours(RETR_TREE,
//   CHAIN_APPROX_SIMPLE); per contour m = cv.moments(contour); centre = int(m10/m00), int(m01/m00)
//   when m00 != 0.
//
// cv.moments of a traced pixel contour is the Green's-theorem polygon moment of the
// polygon through the boundary pixel CENTRES.  For a solid (hole-free) 8-connected
// component that polygon decomposes exactly over 2x2 blocks of pixel centres
// (SURVEY.md §8(a1), Appendix A.1): with n set corners
//     n == 4 : full unit cell      2*area += 2 ; 6*Mx += 6x+3 ; 6*My += 6y+3
//     n == 3 : half cell triangle  2*area += 1 ; 6*Mx += sum x(corners) ; 6*My += sum y(corners)
// so a00, a10, a01 of cv.moments are the integers A2, SX6, SY6 accumulated here, and
// m00 = A2*0.5, m10 = SX6*(1/6), centre = int(m10/m00) is reproduced bit for bit.
//
// Two kernels:
//   k_threshold_segments  HBM-bound stream: every thread loads 16 pixels with one 128-bit
//        load, thresholds them with 3 SWAR integer ops per 4 pixels and appends the rare
//        non-empty 16-bit segment masks to a per-image list (algorithmic bytes: W*H per image,
//        read exactly once; writes are a few hundred bytes per image).
//   k_blob_reduce         one CTA per image on the sparse list only: bitonic sort (raster
//        order), run extraction, union-find over runs in shared memory, per-run cell moments,
//        ranked output in cv.findContours order (descending raster position of first pixel).
#include "common.cuh"

#define SEG_PAD 0xFFFFFFFFu

// ---------------------------------------------------------------------------------------------
// threshold helpers
// ---------------------------------------------------------------------------------------------
struct ThreshConst { uint32_t addc; uint32_t use_and; };   // see swar_gt()

// bit 7 of every byte of the result is (byte > threshold).  T1 = threshold+1 in 1..255:
//   T1 <= 128:  b >= T1  <=>  high bit set  OR  low7 + (128-T1) carries into bit 7
//   T1 >  128:  b >= T1  <=>  high bit set  AND low7 + (256-T1) carries into bit 7
__device__ __forceinline__ uint32_t swar_gt(uint32_t w, ThreshConst tc) {
    uint32_t s = (w & 0x7f7f7f7fu) + tc.addc;
    return tc.use_and ? (s & w) : (s | w);
}
// gathers bit 7 of the four bytes into a nibble (byte 0 -> bit 0)
__device__ __forceinline__ uint32_t nibble_of(uint32_t hi) {
    return ((((hi >> 7) & 0x01010101u) * 0x00204081u) >> 21) & 0xFu;
}

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ void append_segment(uint32_t* seg_count, uint32_t* seg_list, int max_segments,
                                               long long seg_global, int seg_per_image, uint32_t mask16) {
    const int img = (int)(seg_global / seg_per_image);
    const uint32_t pos = (uint32_t)(seg_global - (long long)img * seg_per_image);
    const uint32_t slot = atomicAdd(&seg_count[img], 1u);
    if (slot < (uint32_t)max_segments)
        seg_list[(size_t)img * max_segments + slot] = (pos << 16) | mask16;
}

// 1-channel stream.  UNROLL independent 128-bit loads per thread are issued before any is used.
template <int UNROLL>
__global__ void __launch_bounds__(256)
k_threshold_segments_c1(const uint4* __restrict__ frames, long long n_seg, int seg_per_image,
                        int max_segments, ThreshConst tc, uint32_t* __restrict__ seg_count,
                        uint32_t* __restrict__ seg_list) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; s < n_seg; s += stride * UNROLL) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long si = s + stride * u;
            v[u] = (si < n_seg) ? ldg_stream(frames + si) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t h0 = swar_gt(v[u].x, tc), h1 = swar_gt(v[u].y, tc);
            const uint32_t h2 = swar_gt(v[u].z, tc), h3 = swar_gt(v[u].w, tc);
            if (((h0 | h1 | h2 | h3) & 0x80808080u) == 0) continue;     // the overwhelmingly common case
            const long long si = s + stride * u;
            if (si >= n_seg) continue;
            const uint32_t m = nibble_of(h0) | (nibble_of(h1) << 4) | (nibble_of(h2) << 8) | (nibble_of(h3) << 12);
            append_segment(seg_count, seg_list, max_segments, si, seg_per_image, m);
        }
    }
}

// 3-channel interleaved stream (the layout _find_dot receives).  grey as cv.cvtColor(RGB2GRAY)
// computes it for 8-bit data: (c0*9798 + c1*19235 + c2*3735 + 16384) >> 15  (verified against cv2 4.13).
__global__ void __launch_bounds__(256)
k_threshold_segments_c3(const uint4* __restrict__ frames, long long n_seg, int seg_per_image,
                        int max_segments, int threshold, uint32_t* __restrict__ seg_count,
                        uint32_t* __restrict__ seg_list) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += stride) {
        uint32_t w[12];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const uint4 v = ldg_stream(frames + s * 3 + q);
            w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
        }
        uint32_t any = 0;
#pragma unroll
        for (int q = 0; q < 12; ++q) any |= w[q];
        if (any == 0) continue;
        uint32_t m = 0;
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const int o = 3 * p;
            const uint32_t c0 = (w[o >> 2] >> ((o & 3) * 8)) & 0xffu;
            const uint32_t c1 = (w[(o + 1) >> 2] >> (((o + 1) & 3) * 8)) & 0xffu;
            const uint32_t c2 = (w[(o + 2) >> 2] >> (((o + 2) & 3) * 8)) & 0xffu;
            const int grey = (int)((c0 * 9798u + c1 * 19235u + c2 * 3735u + 16384u) >> 15);
            m |= (grey > threshold ? 1u : 0u) << p;
        }
        if (m) append_segment(seg_count, seg_list, max_segments, s, seg_per_image, m);
    }
}

// ---------------------------------------------------------------------------------------------
// sparse per-image reduction
// ---------------------------------------------------------------------------------------------
// NT == 32: the group is one warp (several images per CTA, warp-level synchronisation only);
// NT  > 32: the group is the whole CTA.
template <int NT>
__device__ __forceinline__ void gsync() {
    if (NT == 32) __syncwarp(); else __syncthreads();
}

template <int NT>
__device__ __forceinline__ unsigned block_scan_excl(unsigned v, unsigned& total, unsigned* wsum) {
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (unsigned)o) x += y;
    }
    if (NT == 32) {
        total = __shfl_sync(0xffffffffu, x, 31);
        return x - v;
    }
    if (lane == 31) wsum[wid] = x;
    __syncthreads();
    unsigned base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
        const unsigned sw = wsum[w];
        if ((unsigned)w < wid) base += sw;
        tot += sw;
    }
    __syncthreads();
    total = tot;
    return base + x - v;
}

__device__ __forceinline__ int seg_find(const uint32_t* seg, int n, uint32_t pos) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((seg[mid] >> 16) < pos) lo = mid + 1; else hi = mid;
    }
    return (lo < n && (seg[lo] >> 16) == pos) ? lo : -1;
}

__device__ __forceinline__ unsigned uf_find(volatile unsigned* parent, unsigned x) {
    unsigned p;
    while ((p = parent[x]) != x) x = p;
    return x;
}
// lock-free union keeping the smaller index as representative (root == first run in raster order)
__device__ __forceinline__ void uf_unite(unsigned* parent, unsigned a, unsigned b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const unsigned tmp = a; a = b; b = tmp; }
        const unsigned old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;
    }
}
__device__ __forceinline__ unsigned run_starts(unsigned m) { return m & ~(m << 1) & 0xffffu; }
// sum of the indices of the set bits (branch-free: weight 2^k times the bits whose index has bit k)
__device__ __forceinline__ int bit_index_sum(unsigned m) {
    return __popc(m & 0xAAAAAAAAu) + 2 * __popc(m & 0xCCCCCCCCu) + 4 * __popc(m & 0xF0F0F0F0u) +
           8 * __popc(m & 0xFF00FF00u) + 16 * __popc(m & 0xFFFF0000u);
}
__device__ __forceinline__ int seg_lower_bound(const uint32_t* seg, int n, uint32_t pos) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((seg[mid] >> 16) < pos) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct BlobSmem {
    uint32_t* seg;        // [E]  sorted (pos<<16)|mask
    unsigned* parent;     // [E]  union-find over runs
    uint16_t* base;       // [E]  first run id of segment i
    uint16_t* node_seg;   // [E]
    uint16_t* node_bits;  // [E]
    uint16_t* rank;       // [E]  blob index of a root run
    unsigned long long* acc;   // [MOCAP_ACC_CAP][4]  A2, SX6, SY6, npix
    unsigned* wsum;       // [32]
};
size_t blob_reduce_smem_bytes(int E) {
    return (size_t)E * (4 + 4 + 2 + 2 + 2 + 2) + (size_t)MOCAP_ACC_CAP * 32 + 32 * 4;
}
__device__ __forceinline__ BlobSmem carve_blob_smem(unsigned char* raw, int E) {
    BlobSmem s;
    s.acc = reinterpret_cast<unsigned long long*>(raw);      raw += (size_t)MOCAP_ACC_CAP * 32;
    s.seg = reinterpret_cast<uint32_t*>(raw);                raw += (size_t)E * 4;
    s.parent = reinterpret_cast<unsigned*>(raw);             raw += (size_t)E * 4;
    s.wsum = reinterpret_cast<unsigned*>(raw);               raw += 32 * 4;
    s.base = reinterpret_cast<uint16_t*>(raw);               raw += (size_t)E * 2;
    s.node_seg = reinterpret_cast<uint16_t*>(raw);           raw += (size_t)E * 2;
    s.node_bits = reinterpret_cast<uint16_t*>(raw);          raw += (size_t)E * 2;
    s.rank = reinterpret_cast<uint16_t*>(raw);
    return s;
}

// Block-wide: the n segments in sm.seg[0..n) (unsorted) -> blobs of one image.
template <bool WIDE>
__device__ __forceinline__ unsigned long long acc_get(const unsigned long long* acc, unsigned idx) {
    return WIDE ? acc[idx] : (unsigned long long)reinterpret_cast<const unsigned*>(acc)[idx];
}

// Returns false (group-uniform) without writing anything when STRICT and a capacity (runs > E,
// blobs > ACC) is exceeded: the caller then hands the image to the full-size kernel.
template <int NT, bool STRICT, bool WIDE>
__device__ bool blob_reduce(BlobSmem sm, int n, int E, int ACC, int W, int H, int max_blobs,
                            int32_t* __restrict__ out_xy, int32_t* __restrict__ out_n,
                            int64_t* __restrict__ out_mom, int32_t* __restrict__ out_flags, int flags_in) {
    const int tid = threadIdx.x % NT;
    const int SPR = W / MOCAP_SEG_PX;     // segments per row
    int flags = flags_in;

    // ---- 1. raster order of (pos<<16 | mask).  Warp groups (n <= 128): rank sort -- every element
    //         counts the smaller ones (keys are unique), no barriers; CTA groups: bitonic sort.
    if (NT == 32) {
        uint32_t* tmp = reinterpret_cast<uint32_t*>(sm.parent);     // free until step 2
        for (int i = tid; i < n; i += NT) {
            const uint32_t e = sm.seg[i];
            int rank = 0;
            for (int j = 0; j < n; ++j) rank += (sm.seg[j] < e) ? 1 : 0;
            tmp[rank] = e;
        }
        gsync<NT>();
        for (int i = tid; i < n; i += NT) sm.seg[i] = tmp[i];
        gsync<NT>();
    } else {
        int n2 = 1;
        while (n2 < n) n2 <<= 1;
        for (int i = n + tid; i < n2; i += NT) sm.seg[i] = SEG_PAD;
        gsync<NT>();
        for (int k = 2; k <= n2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < n2; i += NT) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const uint32_t a = sm.seg[i], b = sm.seg[ixj];
                        const bool up = ((i & k) == 0);
                        if ((a > b) == up) { sm.seg[i] = b; sm.seg[ixj] = a; }
                    }
                }
                gsync<NT>();
            }
        }
    }

    // ---- 2. runs (maximal horizontal strings of set pixels inside one 16-px segment) become nodes
    unsigned n_runs = 0;
    {
        unsigned carry = 0;
        for (int i0 = 0; i0 < n; i0 += NT) {
            const int i = i0 + tid;
            const unsigned cnt = (i < n) ? __popc(run_starts(sm.seg[i] & 0xffffu)) : 0u;
            unsigned tot;
            const unsigned ex = block_scan_excl<NT>(cnt, tot, sm.wsum);
            if (i < n) sm.base[i] = (uint16_t)min(carry + ex, 0xffffu);
            carry += tot;
        }
        n_runs = carry;
    }
    if (n_runs > (unsigned)E) {          // cannot label: report and emit nothing (group-uniform)
        if (STRICT) return false;
        if (tid == 0) { *out_n = 0; if (out_flags) *out_flags = flags | MOCAP_F_SEGMENTS; }
        return true;
    }
    for (int i = tid; i < n; i += NT) {
        unsigned s = sm.seg[i] & 0xffffu;
        unsigned id = sm.base[i];
        while (s) {
            const unsigned b = s & (0u - s);
            const unsigned run = s & ~(s + b);
            s &= ~run;
            sm.node_seg[id] = (uint16_t)i;
            sm.node_bits[id] = (uint16_t)run;
            sm.parent[id] = id;
            ++id;
        }
    }
    gsync<NT>();

    // ---- 3. 8-connectivity unions: right neighbour across the segment boundary, and the row above
    for (unsigned id = tid; id < n_runs; id += NT) {
        const int i = sm.node_seg[id];
        const unsigned rb = sm.node_bits[id];
        const uint32_t p = sm.seg[i] >> 16;
        const int y = p / SPR, sc = p - y * SPR;
        if ((rb & 0x8000u) && sc + 1 < SPR && i + 1 < n) {
            const uint32_t e2 = sm.seg[i + 1];
            if ((e2 >> 16) == p + 1 && (e2 & 1u)) uf_unite(sm.parent, id, sm.base[i + 1]);
        }
        if (y > 0) {                                               // the <= 3 segments of row y-1 that touch this run
            const unsigned ext = (rb | (rb << 1) | (rb >> 1)) & 0xffffu;
            const uint32_t q = p - SPR;
            const int lo = seg_lower_bound(sm.seg, n, sc > 0 ? q - 1 : q);
            for (int j = lo; j < n && j < lo + 3; ++j) {
                const uint32_t ej = sm.seg[j], pj = ej >> 16;
                if (pj > q + 1) break;
                const unsigned mm = ej & 0xffffu;
                if (pj == q) {
                    unsigned sbits = mm, r = 0;
                    while (sbits) {
                        const unsigned b = sbits & (0u - sbits);
                        const unsigned run = sbits & ~(sbits + b);
                        sbits &= ~run;
                        if (run & ext) uf_unite(sm.parent, id, sm.base[j] + r);
                        ++r;
                    }
                } else if (pj + 1 == q) {                           // left neighbour segment (only searched when sc > 0)
                    if ((rb & 1u) && (mm & 0x8000u)) uf_unite(sm.parent, id, sm.base[j] + __popc(run_starts(mm)) - 1);
                } else if (sc + 1 < SPR) {                          // pj == q + 1: right neighbour segment
                    if ((rb & 0x8000u) && (mm & 1u)) uf_unite(sm.parent, id, sm.base[j]);
                }
            }
        }
    }
    gsync<NT>();
    // flatten (two phases so that nobody chases a pointer that is being rewritten;
    // sm.rank is free until step 4 and run ids fit 16 bits because n_runs <= E <= 4096)
    for (unsigned id = tid; id < n_runs; id += NT) sm.rank[id] = (uint16_t)uf_find(sm.parent, id);
    gsync<NT>();
    for (unsigned id = tid; id < n_runs; id += NT) sm.parent[id] = sm.rank[id];
    gsync<NT>();

    // ---- 4. rank the roots (ascending run id == ascending raster position of the blob's first pixel)
    unsigned n_blobs = 0;
    {
        unsigned carry = 0;
        for (unsigned i0 = 0; i0 < n_runs; i0 += NT) {
            const unsigned id = i0 + tid;
            const unsigned is_root = (id < n_runs && sm.parent[id] == id) ? 1u : 0u;
            unsigned tot;
            const unsigned ex = block_scan_excl<NT>(is_root, tot, sm.wsum);
            if (is_root) sm.rank[id] = (uint16_t)min(carry + ex, 0xffffu);
            carry += tot;
        }
        n_blobs = carry;
    }
    if (n_blobs > (unsigned)ACC) {
        if (STRICT) return false;
        flags |= MOCAP_F_BLOBS;
    }
    const unsigned nb = min(n_blobs, (unsigned)ACC);
    for (unsigned k = tid; k < nb * 4; k += NT) { if (WIDE) sm.acc[k] = 0ull; else reinterpret_cast<unsigned*>(sm.acc)[k] = 0u; }
    gsync<NT>();

    // ---- 5. per-run share of the 2x2-cell moments.  A cell is owned by the run holding its
    //         top-left corner, or its top-right corner when the top-left pixel is clear.
    for (unsigned id = tid; id < n_runs; id += NT) {
        const unsigned blob = sm.rank[sm.parent[id]];
        if (blob >= (unsigned)ACC) continue;
        const int i = sm.node_seg[id];
        const unsigned rb = sm.node_bits[id];
        const uint32_t e = sm.seg[i];
        const uint32_t p = e >> 16;
        const unsigned m = e & 0xffffu;
        const int y = p / SPR, sc = p - y * SPR;
        // 18-bit windows over columns 16*sc-1 .. 16*sc+16 of rows y (T) and y+1 (Bw)
        unsigned T = m << 1;
        if (sc > 0 && i > 0 && (sm.seg[i - 1] >> 16) == p - 1) T |= (sm.seg[i - 1] >> 15) & 1u;
        if (sc + 1 < SPR && i + 1 < n && (sm.seg[i + 1] >> 16) == p + 1) T |= (sm.seg[i + 1] & 1u) << 17;
        unsigned Bw = 0;
        if (y + 1 < H) {
            const uint32_t q = p + SPR;
            const int lo = seg_lower_bound(sm.seg, n, sc > 0 ? q - 1 : q);
            for (int j = lo; j < n && j < lo + 3; ++j) {
                const uint32_t ej = sm.seg[j], pj = ej >> 16;
                if (pj > q + 1) break;
                if (pj == q) Bw |= (ej & 0xffffu) << 1;
                else if (pj + 1 == q) Bw |= (ej >> 15) & 1u;        // only searched when sc > 0
                else if (sc + 1 < SPR) Bw |= (ej & 1u) << 17;
            }
        }
        const unsigned Rw = rb << 1;
        const unsigned T1 = T >> 1, B1 = Bw >> 1;
        const unsigned own = (Rw | (~T & (Rw >> 1))) & 0x1ffffu;
        const unsigned full = T & T1 & Bw & B1 & own;
        const unsigned mtl = ~T & T1 & Bw & B1 & own;      // triangle, top-left corner missing
        const unsigned mtr = T & ~T1 & Bw & B1 & own;
        const unsigned mbl = T & T1 & ~Bw & B1 & own;
        const unsigned mbr = T & T1 & Bw & ~B1 & own;
        const int nf = __popc(full), ntl = __popc(mtl), ntr = __popc(mtr), nbl = __popc(mbl), nbr = __popc(mbr);
        const long long x0 = 16ll * sc - 1;               // column of window bit 0
        const long long sxf = x0 * nf + bit_index_sum(full);
        const long long sx_tl = x0 * ntl + bit_index_sum(mtl), sx_tr = x0 * ntr + bit_index_sum(mtr);
        const long long sx_bl = x0 * nbl + bit_index_sum(mbl), sx_br = x0 * nbr + bit_index_sum(mbr);
        const long long a2 = 2ll * nf + ntl + ntr + nbl + nbr;
        const long long sx6 = 6 * sxf + 3ll * nf + 3 * (sx_tl + sx_tr + sx_bl + sx_br) + 2ll * (ntl + nbl) + (ntr + nbr);
        const long long sy6 = (6ll * y + 3) * nf + (3ll * y + 2) * (ntl + ntr) + (3ll * y + 1) * (nbl + nbr);
        if (WIDE) {
            unsigned long long* a = sm.acc + 4 * blob;
            if (a2) {
                atomicAdd(a + 0, (unsigned long long)a2);
                atomicAdd(a + 1, (unsigned long long)sx6);
                atomicAdd(a + 2, (unsigned long long)sy6);
            }
            atomicAdd(a + 3, (unsigned long long)__popc(rb));
        } else {                                                   // 6*max(W,H)*W*H < 2^32: native 32-bit shared atomics
            unsigned* a = reinterpret_cast<unsigned*>(sm.acc) + 4 * blob;
            if (a2) {
                atomicAdd(a + 0, (unsigned)a2);
                atomicAdd(a + 1, (unsigned)sx6);
                atomicAdd(a + 2, (unsigned)sy6);
            }
            atomicAdd(a + 3, (unsigned)__popc(rb));
        }
    }
    gsync<NT>();

    // ---- 6. keep blobs with non-zero polygon area (helpers.py:153), emit in reverse raster order
    unsigned n_keep = 0;
    {
        unsigned carry = 0;
        for (unsigned k0 = 0; k0 < nb; k0 += NT) {      // pass 1: count
            const unsigned k = k0 + tid;
            const unsigned keep = (k < nb && acc_get<WIDE>(sm.acc, 4 * k) != 0ull) ? 1u : 0u;
            unsigned tot;
            block_scan_excl<NT>(keep, tot, sm.wsum);
            carry += tot;
        }
        n_keep = carry;
        carry = 0;
        for (unsigned k0 = 0; k0 < nb; k0 += NT) {      // pass 2: place
            const unsigned k = k0 + tid;
            const unsigned keep = (k < nb && acc_get<WIDE>(sm.acc, 4 * k) != 0ull) ? 1u : 0u;
            unsigned tot;
            const unsigned ex = block_scan_excl<NT>(keep, tot, sm.wsum);
            if (keep) {
                const unsigned o = n_keep - 1 - (carry + ex);
                if (o < (unsigned)max_blobs) {
                    const unsigned long long A2 = acc_get<WIDE>(sm.acc, 4 * k), SX6 = acc_get<WIDE>(sm.acc, 4 * k + 1), SY6 = acc_get<WIDE>(sm.acc, 4 * k + 2);
                    const double m00 = (double)A2 * 0.5;                       // cv.moments: a00 * 0.5
                    const double m10 = (double)SX6 * 0.16666666666666666;      //             a10 * (1/6)
                    const double m01 = (double)SY6 * 0.16666666666666666;
                    out_xy[2 * o + 0] = (int)(m10 / m00);                       // int(m10/m00), helpers.py:154
                    out_xy[2 * o + 1] = (int)(m01 / m00);
                    if (out_mom) {
                        out_mom[4 * o + 0] = (int64_t)A2; out_mom[4 * o + 1] = (int64_t)SX6;
                        out_mom[4 * o + 2] = (int64_t)SY6; out_mom[4 * o + 3] = (int64_t)acc_get<WIDE>(sm.acc, 4 * k + 3);
                    }
                }
            }
            carry += tot;
        }
    }
    if (tid == 0) {
        if (n_keep > (unsigned)max_blobs) flags |= MOCAP_F_BLOBS;
        *out_n = (int)min(n_keep, (unsigned)max_blobs);
        if (out_flags) *out_flags = flags;
    }
    return true;
}

// Sparse reduction, common case: one WARP per image (warp-level synchronisation only).  Images
// with more than WE segments / runs or more than WACC blobs are appended to a worklist for the
// full-size kernel below.  Shared memory per warp is a fixed small slab.
#define BLOB_WE   128     // segments (and runs) a warp handles
#define BLOB_WACC 64      // blobs a warp accumulates
struct WarpSlab {
    unsigned long long acc[BLOB_WACC * 4];
    uint32_t seg[BLOB_WE];
    unsigned parent[BLOB_WE];
    uint16_t base[BLOB_WE], node_seg[BLOB_WE], node_bits[BLOB_WE], rank[BLOB_WE];
};

template <int WPB, bool WIDE>
__global__ void __launch_bounds__(WPB * 32)
k_blob_reduce_warp(uint32_t* __restrict__ seg_count, const uint32_t* __restrict__ seg_list, int n_images, int E,
                   int W, int H, int max_blobs, int32_t* __restrict__ blob_xy, int32_t* __restrict__ blob_n,
                   int64_t* __restrict__ blob_mom, int32_t* __restrict__ img_flags,
                   uint32_t* __restrict__ worklist, uint32_t* __restrict__ work_count) {
    __shared__ WarpSlab slabs[WPB];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x * WPB + warp;
    if (img >= n_images) return;
    const unsigned cnt = seg_count[img];
    int32_t* ofl = img_flags ? img_flags + img : nullptr;
    if (cnt == 0) {
        if (lane == 0) { blob_n[img] = 0; if (ofl) *ofl = 0; }
        return;
    }
    bool ok = cnt <= BLOB_WE;
    if (ok) {
        WarpSlab& sl = slabs[warp];
        BlobSmem sm;
        sm.seg = sl.seg; sm.parent = sl.parent; sm.base = sl.base; sm.node_seg = sl.node_seg;
        sm.node_bits = sl.node_bits; sm.rank = sl.rank; sm.acc = sl.acc; sm.wsum = nullptr;
        const uint32_t* src = seg_list + (size_t)img * E;
        for (int i = lane; i < (int)cnt; i += 32) sm.seg[i] = src[i];
        __syncwarp();
        ok = blob_reduce<32, true, WIDE>(sm, (int)cnt, BLOB_WE, BLOB_WACC, W, H, max_blobs, blob_xy + (size_t)img * max_blobs * 2,
                                   blob_n + img, blob_mom ? blob_mom + (size_t)img * max_blobs * 4 : nullptr, ofl, 0);
    }
    if (lane == 0) {
        if (ok) seg_count[img] = 0;                             // self-cleaning: ready for the next batch
        else worklist[atomicAdd(work_count, 1u)] = (uint32_t)img;
    }
}

// Full-size reduction for the images the warp kernel deferred: persistent CTAs walk the worklist.
template <int NT, bool WIDE>
__global__ void __launch_bounds__(NT)
k_blob_reduce(uint32_t* __restrict__ seg_count, const uint32_t* __restrict__ seg_list, int E, int W, int H,
              int max_blobs, int32_t* __restrict__ blob_xy, int32_t* __restrict__ blob_n,
              int64_t* __restrict__ blob_mom, int32_t* __restrict__ img_flags,
              const uint32_t* __restrict__ worklist, uint32_t* __restrict__ work_count, uint32_t* __restrict__ done_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BlobSmem sm = carve_blob_smem(smem_raw, E);
    const unsigned n_work = *work_count;
    for (unsigned w = blockIdx.x; w < n_work; w += gridDim.x) {
        const int img = (int)worklist[w];
        const unsigned cnt = seg_count[img];
        __syncthreads();
        if (threadIdx.x == 0) seg_count[img] = 0;
        int flags = 0;
        int n = (int)cnt;
        if (cnt > (unsigned)E) { flags |= MOCAP_F_SEGMENTS; n = 0; }
        int32_t* ofl = img_flags ? img_flags + img : nullptr;
        if (n == 0) {
            if (threadIdx.x == 0) { blob_n[img] = 0; if (ofl) *ofl = flags; }
            continue;
        }
        const uint32_t* src = seg_list + (size_t)img * E;
        for (int i = threadIdx.x; i < n; i += NT) sm.seg[i] = src[i];
        __syncthreads();
        blob_reduce<NT, false, WIDE>(sm, n, E, MOCAP_ACC_CAP, W, H, max_blobs, blob_xy + (size_t)img * max_blobs * 2, blob_n + img,
                               blob_mom ? blob_mom + (size_t)img * max_blobs * 4 : nullptr, ofl, flags);
        __syncthreads();
    }
    if (threadIdx.x == 0) {                                     // the last CTA to finish re-arms the worklist
        __threadfence();
        if (atomicAdd(done_count, 1u) == gridDim.x - 1) { *work_count = 0; *done_count = 0; }
    }
}

// ---------------------------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------------------------
int launch_detect(mocap_ctx* ctx, const uint8_t* frames, int n_images, int channels, int threshold,
                  int32_t* blob_xy, int32_t* blob_n, int64_t* blob_mom, int32_t* img_flags) {
    const mocap_config& c = ctx->cfg;
    const int seg_per_image = c.width * c.height / MOCAP_SEG_PX;
    const long long n_seg = (long long)n_images * seg_per_image;
    const int E = c.max_segments;
    if (n_images <= 0) return MOCAP_OK;

    if (ctx->timing_on) {
        if (ctx->tim_used == 64) { const int st = timing_flush(ctx); if (st) return st; }
        CUDA_TRY(ctx, cudaEventRecord(ctx->tim_ev[2 * ctx->tim_used], ctx->stream));
    }
    const int threads = 256;
    if (channels == 1) {
        constexpr int UNROLL = 8;
        ThreshConst tc;
        if (threshold < 0) { tc.addc = 0x80808080u; tc.use_and = 0; }            // everything passes
        else if (threshold >= 255) { tc.addc = 0; tc.use_and = 1; }              // nothing passes
        else {
            const uint32_t T1 = (uint32_t)threshold + 1u;
            tc.use_and = T1 > 128 ? 1u : 0u;
            tc.addc = (T1 > 128 ? 256u - T1 : 128u - T1) * 0x01010101u;
        }
        long long want = (n_seg + (long long)threads * UNROLL - 1) / ((long long)threads * UNROLL);
        const long long cap = (long long)ctx->num_sms * 8;      // 8 CTAs of 256 threads fill an SM
        const int grid = (int)(want < cap ? want : cap);
        k_threshold_segments_c1<UNROLL><<<grid, threads, 0, ctx->stream>>>(
            reinterpret_cast<const uint4*>(frames), n_seg, seg_per_image, E, tc, ctx->d_seg_count, ctx->d_seg_list);
    } else {
        long long want = (n_seg + threads - 1) / threads;
        const long long cap = (long long)ctx->num_sms * 8;
        const int grid = (int)(want < cap ? want : cap);
        k_threshold_segments_c3<<<grid, threads, 0, ctx->stream>>>(
            reinterpret_cast<const uint4*>(frames), n_seg, seg_per_image, E, threshold, ctx->d_seg_count, ctx->d_seg_list);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    if (ctx->timing_on) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->tim_ev[2 * ctx->tim_used + 1], ctx->stream));
        ctx->tim_used += 1;
    }
    constexpr int WPB = 8;
    constexpr int NT = 128;
    const size_t smem = blob_reduce_smem_bytes(E);
    const int grid2 = n_images < ctx->num_sms ? n_images : ctx->num_sms;
    const long long mx = c.width > c.height ? c.width : c.height;
    const bool wide = 6ll * mx * c.width * c.height >= (1ll << 32);      // moment sums may exceed 32 bits
    if (wide) {
        k_blob_reduce_warp<WPB, true><<<(n_images + WPB - 1) / WPB, WPB * 32, 0, ctx->stream>>>(
            ctx->d_seg_count, ctx->d_seg_list, n_images, E, c.width, c.height, c.max_blobs, blob_xy, blob_n, blob_mom, img_flags,
            ctx->d_worklist, ctx->d_work_count);
        CUDA_TRY(ctx, cudaGetLastError());
        k_blob_reduce<NT, true><<<grid2, NT, smem, ctx->stream>>>(ctx->d_seg_count, ctx->d_seg_list, E, c.width, c.height,
                                                               c.max_blobs, blob_xy, blob_n, blob_mom, img_flags,
                                                               ctx->d_worklist, ctx->d_work_count, ctx->d_work_count + 1);
    } else {
        k_blob_reduce_warp<WPB, false><<<(n_images + WPB - 1) / WPB, WPB * 32, 0, ctx->stream>>>(
            ctx->d_seg_count, ctx->d_seg_list, n_images, E, c.width, c.height, c.max_blobs, blob_xy, blob_n, blob_mom, img_flags,
            ctx->d_worklist, ctx->d_work_count);
        CUDA_TRY(ctx, cudaGetLastError());
        k_blob_reduce<NT, false><<<grid2, NT, smem, ctx->stream>>>(ctx->d_seg_count, ctx->d_seg_list, E, c.width, c.height,
                                                                c.max_blobs, blob_xy, blob_n, blob_mom, img_flags,
                                                                ctx->d_worklist, ctx->d_work_count, ctx->d_work_count + 1);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 1;
    ctx->launches += 2;
    return MOCAP_OK;
}

int timing_flush(mocap_ctx* ctx) {
    for (int i = 0; i < ctx->tim_used; ++i) {
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->tim_ev[2 * i + 1]));
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->tim_ev[2 * i], ctx->tim_ev[2 * i + 1]));
        ctx->detect_ms_sum += ms;
        ctx->detect_ms_n += 1;
    }
    ctx->tim_used = 0;
    return MOCAP_OK;
}

int blob_kernels_init(mocap_ctx* ctx) {
    const size_t smem = blob_reduce_smem_bytes(ctx->cfg.max_segments);
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_blob_reduce<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_blob_reduce<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return MOCAP_OK;
}
