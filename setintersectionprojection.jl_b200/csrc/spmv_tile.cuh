// spmv_tile.cuh — tiled, TMA-staged 7-point CDS SpMV for 3-D grids (sm_100a).
//
// Replaces CDS_MVp_MT.jl:9-25 / Ax_CDS_MT (argmin_x.jl:72-78) [+ dot(p,Ap), cg.jl:88] and, in MODE 2, the CG
// prologue (argmin_x.jl:33-37 + cg.jl:47-76) for matrices whose offsets are a subset of {0, +-1, +-n0, +-n0*n1}
// on an (n0, n1, n2) grid — every Q the reference assembles from get_TD_operator.jl on a 3-D grid.
//
// Design (2.5-D plane sweep):
//   * a CTA owns a tile of TJ whole grid lines (all i, j0 <= j < j0+TJ) and marches along z through a chunk of
//     planes [ka, kb).  Whole lines make every plane-tile ONE contiguous range of the vector, so a plane-tile
//     (plus one halo line on each side, which also holds the +-1 neighbours of the line ends) is fetched by a single
//     1-D bulk copy (cp.async.bulk.shared::cluster.global, SASS UBLKCP) into a ring of NS shared-memory stages,
//     completion signalled on an mbarrier; NS-2 plane-tiles are in flight while one is being consumed.  There is no
//     block-wide barrier inside the sweep and no thread ever waits for a stage to drain: every warp counts itself
//     out of a stage when it is done with it, and the LAST warp to leave refills that stage at once.
//   * every thread owns ONE 16-byte column group of up to TileItems<T>::n lines of the tile: the z-1 and centre
//     vectors rotate through registers, the z+1 vector and the +-n0 / +-1 neighbours come from shared memory
//     (3 LDS.128 + 2 LDS.32 per group instead of 7 cached global loads), y leaves with one 16-byte store.  The
//     column is fixed per thread, so the stencil-class coefficients of a group live in registers and change only
//     on the first / last line or plane of the grid; planes strictly inside the grid run a predicate-free body.
//   * the accumulation order of the seven terms is a compile-time permutation (kTileOrders: the orders Q_offsets
//     takes for the reference's operator sets); any other order falls back to the generic kernel.
//   * x is read from DRAM once (halo lines are L2 hits: the neighbouring tile's CTA marches in lockstep),
//     y is written once: 2*N*s bytes in class form, (nd+2)*N*s in array form (diagonals streamed with
//     16-byte ld.global.cs).
//   * the multiply-adds are evaluated in the order of Q_offsets, every product rounded separately, exactly as
//     the generic kernel does: results are bit-identical (tests/test_gpu_parity.py).
//   * slabs: the planes below / above the slab come from the local halo planes or, on the peer path, straight
//     from the neighbour's memory over NVLink with the same bulk copy (after the version-flag wait).
#pragma once
#include "common.cuh"

namespace sipb {

#ifndef SIPB_TILE_ITEMS_F32
#define SIPB_TILE_ITEMS_F32 4
#endif
#ifndef SIPB_TILE_MINCTAS
#define SIPB_TILE_MINCTAS 2
#endif
constexpr int kTileMinCtas = SIPB_TILE_MINCTAS;     // resident CTAs per SM the kernel is compiled for
template <typename T> struct TileItems { static constexpr int n = sizeof(T) == 4 ? SIPB_TILE_ITEMS_F32 : 2; };   // (line, column group) pairs per thread
constexpr int kTileMaxStages = 5;
#ifndef SIPB_TILE_LOCKSTEP
#define SIPB_TILE_LOCKSTEP 1
#endif
constexpr bool kTileLockstep = SIPB_TILE_LOCKSTEP != 0;   // 1: one block barrier per plane; 0: last warp out of a stage refills it
constexpr int kTileDiag = 8;        // diagonals handled (7 used by a 3-D stencil)
constexpr int kTileSlots = 7;       // 0 centre, 1 (-1), 2 (+1), 3 (-n0), 4 (+n0), 5 (-P), 6 (+P)
constexpr int kTileNumOrders = 3;
// accumulation orders (slot sequences) with a compiled body: Q_offsets of [identity, TV, ...] = (0,-P,-n0,-1,1,n0,P),
// of [identity, D_z, D_x, D_y] = (0,-P,P,-1,1,-n0,n0), and ascending offsets; a matrix with fewer diagonals matches an
// order when its sequence is the order restricted to the slots it has (the tiled kernel wants all seven)
__host__ __device__ constexpr int tile_order_slot(int ord, int pos) {
  constexpr int t[kTileNumOrders][kTileSlots] = {{0, 5, 3, 1, 2, 4, 6}, {0, 5, 6, 1, 2, 3, 4}, {5, 3, 1, 0, 2, 4, 6}};
  return t[ord][pos];
}

struct TileGeom {
  int ok;
  int TJ, JT, KC, NS;
  int BD;               // threads per CTA (a whole number of warps covering LPP lines)
  int LPP;              // lines covered by one pass of the CTA
  int order;            // index into the compiled accumulation orders (tile_order_slot)
  int dcol[kTileDiag];  // slot -> column of the matrix (diagonal index in accumulation order)
  int gpl;              // 16-byte groups per grid line
  int nloc;             // local planes
  int kofs;             // global index of local plane 0
  int n2g;              // global planes
  int has_lo, has_hi;   // a plane below / above the slab exists in memory (local halo or peer pointer)
  int stage_elems;      // elements per stage
  int code[kTileDiag];  // per diagonal in accumulation order: 0 centre, 1 (-1), 2 (+1), 3 (-n0), 4 (+n0), 5 (-P), 6 (+P)
  int grid;             // CTAs to launch
  size_t smem_bytes;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <typename T>
struct TileInit {       // MODE 2 (CG prologue): r = b - Q x ; p = r ; x_old = x ; sums bb, rr
  const T* b;
  T* r;
  T* p;
  T* x_old;
};

template <typename T> __device__ __forceinline__ void zero_vec(T (&v)[Vec<T>::W]) {
#pragma unroll
  for (int e = 0; e < Vec<T>::W; ++e) v[e] = (T)0;
}

// position in the ring of stages: stage index and the parity of its current use
struct RingPos {
  int stage;
  unsigned parity;
  __device__ __forceinline__ void advance(int ns) {
    if (++stage == ns) { stage = 0; parity ^= 1u; }
  }
};

// One 16-byte group of rows.  xs[slot][e]: the seven neighbour vectors; c0 / cm / cL: class coefficients per slot for the
// first element, the middle elements and the last element of the group (class form); off: row of the group inside its
// plane, `plane_rows`: first row of the plane.
template <typename T, int MODE, bool ARR, int ORD>
__device__ __forceinline__ void tile_group(const SpmvArgs<T>& a, const TileGeom& g, const TileInit<T>& ia, i64 plane_rows,
                                           int off, const T (&xs)[kTileSlots][Vec<T>::W], const T (&c0)[kTileSlots],
                                           const T (&cm)[kTileSlots], const T (&cL)[kTileSlots], double (&dsum)[2]) {
  constexpr int VW = Vec<T>::W;
  T acc[VW];
  zero_vec<T>(acc);
#pragma unroll
  for (int pos = 0; pos < kTileSlots; ++pos) {
    const int s = tile_order_slot(ORD, pos);
    // the first term is 0 + q*x == q*x (up to the sign of a zero result): no addition
    if (ARR) {
      T qv[VW];
      vload_stream<T>(a.R + (i64)g.dcol[s] * a.ld + plane_rows + off, qv);
#pragma unroll
      for (int e = 0; e < VW; ++e) acc[e] = pos == 0 ? qv[e] * xs[s][e] : acc[e] + qv[e] * xs[s][e];
    } else {
#pragma unroll
      for (int e = 0; e < VW; ++e) {
        const T q = e == 0 ? c0[s] : (e == VW - 1 ? cL[s] : cm[s]);
        acc[e] = pos == 0 ? q * xs[s][e] : acc[e] + q * xs[s][e];
      }
    }
  }
  if (MODE == 2) {
    T bv[VW], rv[VW];
    vload_stream<T>(ia.b + plane_rows + off, bv);
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      rv[e] = bv[e] - acc[e];
      dsum[0] += (double)bv[e] * (double)bv[e];
      dsum[1] += (double)rv[e] * (double)rv[e];
    }
    vstore<T>(ia.r + plane_rows + off, rv);
    vstore<T>(ia.p + plane_rows + off, rv);
    if (ia.x_old) vstore<T>(ia.x_old + plane_rows + off, xs[0]);
  } else {
    vstore<T>(a.y + plane_rows + off, acc);
    if (MODE == 1) {
#pragma unroll
      for (int e = 0; e < VW; ++e) dsum[0] += (double)xs[0][e] * (double)acc[e];
    }
  }
}

template <typename T>
__device__ __forceinline__ void tile_load_class(const SpmvArgs<T>& a, const TileGeom& g, const T* tab_s, int cls, int col,
                                                T (&c0)[kTileSlots], T (&cm)[kTileSlots], T (&cL)[kTileSlots]) {
  constexpr int VW = Vec<T>::W;
  const int i0 = col == 0 ? 0 : 1, iL = col + VW == (int)a.gn[0] ? 2 : 1;
#pragma unroll
  for (int s = 0; s < kTileSlots; ++s) {
    c0[s] = tab_s[(cls * 3 + i0) * a.nd + g.dcol[s]];
    cm[s] = tab_s[(cls * 3 + 1) * a.nd + g.dcol[s]];
    cL[s] = tab_s[(cls * 3 + iL) * a.nd + g.dcol[s]];
  }
}

// One plane of a unit.  EDGE = false: the plane lies strictly inside the GLOBAL grid: every neighbour exists (no
// predicates) and the class of a row depends on its line only (jcls: 2 bits per owned line; it differs from
// "interior" only on the first / last line of the grid).  EDGE = true: first / last plane of the grid: neighbours
// outside the vector count as zero (CDS_MVp.jl:14-17 clips the row range instead; the products are then +-0 and
// change nothing).
template <typename T, int MODE, bool ARR, int ORD, bool EDGE>
__device__ __forceinline__ void tile_plane(const SpmvArgs<T>& a, const TileGeom& g, const TileInit<T>& ia, const T* sc,
                                           const T* sp, const T* tab_s, int kl, int j0, int nval, int idx0, int off0,
                                           int col, int line0, unsigned jcls, int& cur_cls, T (&c0)[kTileSlots], T (&cm)[kTileSlots],
                                           T (&cL)[kTileSlots], T (&vkm)[TileItems<T>::n][Vec<T>::W],
                                           T (&vc)[TileItems<T>::n][Vec<T>::W], double (&dsum)[2]) {
  constexpr int VW = Vec<T>::W;
  constexpr int NIT = TileItems<T>::n;
  const int n0 = (int)a.gn[0], n1 = (int)a.gn[1];
  const int lstride = g.LPP * n0;
  const int kg = kl + g.kofs;
  const bool has_km = !EDGE || kg > 0, has_kp = !EDGE || kg < g.n2g - 1;
  const i64 plane_rows = (i64)kl * n0 * n1;
#pragma unroll
  for (int q = 0; q < NIT; ++q) {
    if (q >= nval) break;
    const int idx = idx0 + q * lstride;
    T xs[kTileSlots][VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      xs[0][e] = vc[q][e];
      xs[5][e] = has_km ? vkm[q][e] : (T)0;
    }
    T xm, xp;
    if (EDGE) {
      const int j = j0 + line0 + q * g.LPP;
      const bool has_jm = has_km || j > 0, has_jp = has_kp || j < n1 - 1;
      zero_vec<T>(xs[6]); zero_vec<T>(xs[3]); zero_vec<T>(xs[4]);
      if (has_kp) vload<T>(sp + idx, xs[6]);
      if (has_jm) vload<T>(sc + idx - n0, xs[3]);
      if (has_jp) vload<T>(sc + idx + n0, xs[4]);
      xm = (col > 0 || has_jm) ? sc[idx - 1] : (T)0;
      xp = (col + VW < n0 || has_jp) ? sc[idx + VW] : (T)0;
      if (!ARR) {
        const int kcls = kg == 0 ? 0 : (kg == g.n2g - 1 ? 2 : 1);
        const int cls = 9 + kcls * 3 + (j == 0 ? 0 : (j == n1 - 1 ? 2 : 1));
        if (cls != cur_cls) {
          cur_cls = cls;
          tile_load_class<T>(a, g, tab_s, cls - 9, col, c0, cm, cL);
        }
      }
    } else {
      vload<T>(sp + idx, xs[6]);
      vload<T>(sc + idx - n0, xs[3]);
      vload<T>(sc + idx + n0, xs[4]);
      xm = sc[idx - 1];
      xp = sc[idx + VW];
      if (!ARR) {
        const int cls = 12 + (int)((jcls >> (2 * q)) & 3u);      // kcls = 1
        if (cls != cur_cls) {      // rare: first / last line of the grid
          cur_cls = cls;
          tile_load_class<T>(a, g, tab_s, cls - 9, col, c0, cm, cL);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < VW; ++e) {
      xs[1][e] = e == 0 ? xm : vc[q][e - 1];
      xs[2][e] = e == VW - 1 ? xp : vc[q][e + 1];
    }
    tile_group<T, MODE, ARR, ORD>(a, g, ia, plane_rows, off0 + q * lstride, xs, c0, cm, cL, dsum);
#pragma unroll
    for (int e = 0; e < VW; ++e) {       // rotate the z pipeline
      vkm[q][e] = vc[q][e];
      vc[q][e] = xs[6][e];
    }
  }
}

// MODE 1: y = Q x and the partial sums of dot(x, y) (out_dot == null: the sum is dropped).  MODE 2: CG prologue.
// ARR: the matrix comes from the CDS arrays a.R (else from the stencil-class table a.tab).  ORD: accumulation order.
template <typename T, int MODE, bool ARR, int ORD>
__global__ void __launch_bounds__(kThreads, kTileMinCtas)
    k_spmv_tile(SpmvArgs<T> a, const __grid_constant__ TileGeom g, TileInit<T> ia, RedScratch rs, double* out_dot,
                const int* __restrict__ done_flag, CgState* st, const __grid_constant__ CommDev cd) {
  if (done_flag && *done_flag) return;
  constexpr int VW = Vec<T>::W;
  constexpr int NIT = TileItems<T>::n;
  extern __shared__ __align__(128) unsigned char tile_smem[];
  T* stages = reinterpret_cast<T*>(tile_smem);
  T* tab_s = stages + (size_t)g.NS * g.stage_elems;
  __shared__ __align__(8) unsigned long long full_bar[kTileMaxStages];
  __shared__ unsigned rel_cnt[kTileMaxStages];      // warps that have left the stage
  const int tid = threadIdx.x;
  const int n0 = (int)a.gn[0], n1 = (int)a.gn[1];
  const i64 P = (i64)n0 * n1;
  if (tid == 0) {
    for (int s = 0; s < g.NS; ++s) {
      mbar_init(&full_bar[s], 1u);
      rel_cnt[s] = 0u;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!ARR) {
    for (int q = tid; q < 27 * a.nd; q += blockDim.x) tab_s[q] = a.tab[q];
  }
  __syncthreads();

  // this thread's column group and first line inside a tile — the same for every unit
  const int line0 = tid / g.gpl;
  const int col = (tid - line0 * g.gpl) * VW;
  const bool active = line0 < g.LPP;
  T c0[kTileSlots], cm[kTileSlots], cL[kTileSlots];
  int cur_cls = -1;
#pragma unroll
  for (int s = 0; s < kTileSlots; ++s) { c0[s] = (T)0; cm[s] = (T)0; cL[s] = (T)0; }

  double dsum[2] = {0.0, 0.0};
  const unsigned nwarps = blockDim.x >> 5;
  RingPos cons{0, 0u};         // ring position of load 0 of the current unit
  const int units = g.JT * g.KC;
  bool first_unit = true;
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const int jt = unit % g.JT, kc = unit / g.JT;
    const int j0 = jt * g.TJ;
    const int TJt = min(g.TJ, n1 - j0);
    const int ka = (int)((i64)g.nloc * kc / g.KC), kb = (int)((i64)g.nloc * (kc + 1) / g.KC);
    if (ka == kb) continue;
    if (!first_unit) __syncthreads();               // every warp has left every stage of the previous unit
    first_unit = false;
    if (cd.on && (a.x_lo || a.x_hi)) {
      // peer path: the planes next to the slab are read from the neighbours' memory — wait for their version flag
      if ((ka == 0 && a.x_lo) || (kb == g.nloc && a.x_hi)) {
        p_wait(cd, ka == 0, kb == g.nloc);
        fence_proxy_async();
      }
    }
    const int nq = kb - ka + 2;                     // plane-tiles of this unit: planes ka-1 .. kb
    const i64 tile_off = (i64)(j0 - 1) * n0;        // element offset of stage index 0 inside its plane
    // ---- producer: ONE thread issues the bulk copy of plane-tile q of this unit into ring position `pos` --------
    auto issue = [&](int q, const RingPos& pos) {
      const int kl = ka - 1 + q;
      unsigned long long* bar = &full_bar[pos.stage];
      T* sdst = stages + (size_t)pos.stage * g.stage_elems;
      const bool halo = q >= 1 && q <= nq - 2;
      // element range relative to the start of local plane kl
      const i64 e_lo = halo ? tile_off : (i64)j0 * n0;
      const i64 e_hi = halo ? (i64)(j0 + TJt + 1) * n0 : (i64)(j0 + TJt) * n0;
      // split by plane: the part below plane kl, inside it, above it
      const void* src[3];
      unsigned bytes[3];
      i64 dsto[3];
      unsigned total = 0;
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        const int pl = kl - 1 + part;                 // plane the part lies in
        const i64 lo = max(e_lo, (i64)(part - 1) * P), hi = min(e_hi, (i64)part * P);
        bytes[part] = 0;
        if (hi <= lo) continue;
        const T* base = nullptr;                      // pointer to element 0 of plane pl
        if (pl < 0) {
          if (pl == -1 && g.has_lo) base = a.x_lo ? a.x_lo + (a.n_lo - P) : a.x - P;
        } else if (pl >= g.nloc) {
          if (pl == g.nloc && g.has_hi) base = a.x_hi ? a.x_hi : a.x + (i64)g.nloc * P;
        } else {
          base = a.x + (i64)pl * P;
        }
        if (!base) continue;
        src[part] = base + (lo - (i64)(part - 1) * P);
        dsto[part] = lo - tile_off;
        bytes[part] = (unsigned)((hi - lo) * (i64)sizeof(T));
        total += bytes[part];
      }
      mbar_expect_tx(bar, total);
#pragma unroll
      for (int part = 0; part < 3; ++part)
        if (bytes[part]) bulk_g2s(sdst + dsto[part], src[part], bytes[part], bar);
    };
    // a warp is done with the stage at ring position `pos` that held load q: the last warp out refills it with load
    // q + NS (same stage, next use)
    auto release = [&](int q, const RingPos& pos) {
      if (kTileLockstep) {       // block-wide step: thread 0 refills the stage once every warp has passed the barrier
        __syncthreads();
        if (tid == 0 && q + g.NS < nq) {
          RingPos nxt_use{pos.stage, pos.parity ^ 1u};
          issue(q + g.NS, nxt_use);
        }
        return;
      }
      __syncwarp();
      if ((tid & 31) == 0) {
        const unsigned old = atomicAdd(&rel_cnt[pos.stage], 1u);
        if (old == nwarps - 1u) {
          rel_cnt[pos.stage] = 0u;
          if (q + g.NS < nq) {
            RingPos nxt_use{pos.stage, pos.parity ^ 1u};
            issue(q + g.NS, nxt_use);
          }
        }
      }
    };
    if (tid == 0) {
      RingPos pos = cons;
      for (int q = 0; q < min(g.NS, nq); ++q) {
        issue(q, pos);
        pos.advance(g.NS);
      }
    }

    // per-thread constants of the unit
    int nval = 0;                                   // lines of the tile this thread owns
    if (active)
      for (int q = 0; q < NIT; ++q) nval += (line0 + q * g.LPP < TJt) ? 1 : 0;
    const int idx0 = (line0 + 1) * n0 + col;        // index of the first owned group inside a stage
    const int off0 = (j0 + line0) * n0 + col;       // its row inside a plane
    unsigned jcls = 0u;                             // class of every owned line along j, 2 bits each
    for (int q = 0; q < NIT; ++q) {
      const int j = j0 + line0 + q * g.LPP;
      jcls |= (j == 0 ? 0u : (j == n1 - 1 ? 2u : 1u)) << (2 * q);
    }

    // ---- prologue: z-1 and centre vectors of the first plane ----------------------------------------------
    T vkm[NIT][VW], vc[NIT][VW];
    RingPos cur = cons;                             // ring position of the plane being consumed
    mbar_wait(&full_bar[cur.stage], cur.parity);
    {
      const T* s0 = stages + (size_t)cur.stage * g.stage_elems;
      const bool have_km = (ka - 1 + g.kofs) >= 0 && (ka - 1 >= 0 || g.has_lo);
#pragma unroll
      for (int q = 0; q < NIT; ++q) {
        zero_vec<T>(vkm[q]);
        if (q < nval && have_km) vload<T>(s0 + idx0 + q * g.LPP * n0, vkm[q]);
      }
    }
    release(0, cur);
    cur.advance(g.NS);
    mbar_wait(&full_bar[cur.stage], cur.parity);
    {
      const T* s1 = stages + (size_t)cur.stage * g.stage_elems;
#pragma unroll
      for (int q = 0; q < NIT; ++q) {
        zero_vec<T>(vc[q]);
        if (q < nval) vload<T>(s1 + idx0 + q * g.LPP * n0, vc[q]);
      }
    }
    // ---- plane sweep ---------------------------------------------------------------------------------------
    for (int kl = ka; kl < kb; ++kl) {
      RingPos nxt = cur;
      nxt.advance(g.NS);
      mbar_wait(&full_bar[nxt.stage], nxt.parity);
      const T* sc = stages + (size_t)cur.stage * g.stage_elems;
      const T* sp = stages + (size_t)nxt.stage * g.stage_elems;
      const int kg = kl + g.kofs;
      if (kg > 0 && kg < g.n2g - 1)
        tile_plane<T, MODE, ARR, ORD, false>(a, g, ia, sc, sp, tab_s, kl, j0, nval, idx0, off0, col, line0, jcls, cur_cls, c0,
                                             cm, cL, vkm, vc, dsum);
      else
        tile_plane<T, MODE, ARR, ORD, true>(a, g, ia, sc, sp, tab_s, kl, j0, nval, idx0, off0, col, line0, jcls, cur_cls, c0,
                                            cm, cL, vkm, vc, dsum);
      release(kl - ka + 1, cur);                    // this warp is done with plane kl's stage
      cur = nxt;
    }
    release(nq - 1, cur);                           // the stage of plane kb (read as z+1 by the last step)
    cur.advance(g.NS);
    cons = cur;
  }

  // ---- reductions / publication (same protocol as k_spmv / k_cg_init) --------------------------------------
  if (MODE == 1) {
    double d1[1] = {dsum[0]};
    if (grid_sum<1>(d1, rs)) {
      if (cd.on && out_dot) mail_allreduce<1>(cd, d1, out_dot);      // the global p.Ap is in place when the kernel ends
      else if (threadIdx.x == 0 && out_dot) out_dot[0] = d1[0];
    }
  } else if (MODE == 2) {
    // only CTAs that wrote boundary planes of r need the system-scope fence (the neighbours read those planes)
    bool sys = false;
    if (cd.on) {
      for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
        const int kc = unit / g.JT;
        const int ka = (int)((i64)g.nloc * kc / g.KC), kb = (int)((i64)g.nloc * (kc + 1) / g.KC);
        sys = sys || (ka < kb && (ka == 0 || kb == g.nloc));
      }
    }
    if (grid_sum<2>(dsum, rs, sys)) {
      if (cd.on) {
        mail_publish<2>(cd, dsum);
      } else if (threadIdx.x == 0) {
        st->bb = dsum[0];
        st->rr = dsum[1];
      }
    }
  }
}

}  // namespace sipb
