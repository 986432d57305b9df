// problem.cuh — device-resident least-squares problem shared by problem.cu (linearisation) and solver.cu (LM + band solver).
//
// HBM layout:
//   X        one flat fp64 parameter vector [r3 3n | so3 4n | sens 22 | rho nl]; a second copy XC holds the LM candidate
//   tables   per residual type SoA (ResTable, residuals.cuh), in the caller's (chronological) order
//   H / A    normal equations in 32x32 TILE storage: block column J holds tiles d = 0..T (band, rows 32(J+d)..) followed by
//            RB border tiles (arrow rows: map-time knots + sensor blocks + one extra row carrying the right-hand side);
//            the border x border corner is a small dense matrix C.  Never dense n^2 (SURVEY §5 "long-context").
#pragma once
#include <memory>

#include "common.cuh"
#include "lowering.hpp"

namespace lvi {

constexpr int kTileLog = 5;
constexpr int kTile = 1 << kTileLog;      // 32 x 32 fp64 tiles (8 KB): measured best — the factorisation is bound by its serial
                                          // pivot chain, and the per-column chain cost grows faster than the tile width
constexpr int kTileElems = kTile * kTile;

struct BandSys {
  int nb, nbo;       // band dims, border dims (without the rhs row)
  int NT, T, RB;     // block columns, sub-diagonal tile rows, border tile rows
  int NT0;           // two-sided ordering: block columns [0, NT0) are chain 0, [NT0, NT) chain 1 (NT0 == NT: one chain)
  int n_mid;         // leading border dims that form the separator between the chains (factored as a second-level system)
  int TPC;           // tiles per block column = T + 1 + RB
  int ldc;           // RB * kTile
  double* tiles;     // [NT * TPC * kTileElems], each tile column-major
  double* C;         // [ldc * ldc] column-major, lower
  double* Linv;      // [NT * kTileElems] inverse of each diagonal Cholesky block (filled by the factorisation)
  double* x;         // [NT*kTile + ldc] solution (band part then border part)
  int* fail;         // != 0: Cholesky breakdown
  int* work_i;       // [NT*TPC tile-ready flags | NT back-substitution arrival counters | 2 task counters | 2 chain SM ids] (zeroed per solve)
  unsigned long long* trace;  // optional [NT*TPC*8] per-task timestamps (LVI_TRACE_FACTOR=file), else nullptr
  double* work_d;    // [NT*kTile] partial sums of the back substitution, then [NT*2*kTile] 8-byte mailbox words (value half | flag)
  size_t work_i_count() const { return static_cast<size_t>(NT) * TPC + NT + 2 + 2; }   // + SM ids of the chain CTAs
  // vsum [NT*32] | back-substitution mailbox [NT*64 words]
  size_t work_d_count() const { return static_cast<size_t>(NT) * 3 * kTile; }
  // flagged ("LL") copies: one per tile, W_j and the pre-accumulated first sub-diagonal tile per column; 16 B per value; zeroed ONCE at
  // allocation, valid words carry the epoch of the factorisation that wrote them
  unsigned long long* ll;
  unsigned epoch;
  int pre_shift;     // column slots by which the pre-accumulation tasks are queued ahead of their column (set per launch)
  unsigned char lead[64]; // column slots by which band tile s (distance to the diagonal) is queued ahead of the border tiles: pre_shift for s <= 1, less further out (set per launch)
  // ... followed by the flagged vectors of the back substitution (x_m and u_m, 32 values = 64 words each per column)
  size_t ll_count() const { return (static_cast<size_t>(NT) * TPC + 2 * static_cast<size_t>(NT)) * 2 * kTileElems + static_cast<size_t>(NT) * 4 * kTile; }
};

// address of H(i,j), i >= j, positions in [0, nb+nbo]; position nb+nbo is the rhs row
__device__ __forceinline__ double* band_addr(const BandSys& S, int i, int j) {
  if (j >= S.nb) {
    return S.C + (i - S.nb) + static_cast<size_t>(S.ldc) * (j - S.nb);
  }
  const int J = j >> kTileLog;
  size_t tile;
  int a;
  if (i >= S.nb) { const int b = i - S.nb; tile = static_cast<size_t>(J) * S.TPC + S.T + 1 + (b >> kTileLog); a = b & (kTile - 1); }
  else { tile = static_cast<size_t>(J) * S.TPC + ((i >> kTileLog) - J); a = i & (kTile - 1); }
  return S.tiles + (tile << (2 * kTileLog)) + ((j & (kTile - 1)) << kTileLog) + a;
}

// Schur rows of the inverse depths (lowering.hpp): H_rr (diagonal), H_rx as one dense row per landmark over its coupling slots
struct SchurView {
  int n_rho, base;          // base = nb + nbo: tangent position of the first inverse depth
  const int* row_start;     // [n_landmarks + 1] by landmark id
  const int* row_pos;       // [slots] band/border position of each slot (-1: constant)
  const int* lm_of_rho;     // [n_rho] landmark id of the k-th free inverse depth
  const int* slot2u;        // [slots] index of the slot's position in the landmark's list of DISTINCT positions (-1: constant)
  const int* urow_pos;      // [slots] the distinct positions of landmark l, ascending, at row_start[l] .. row_start[l] + ulen[l]
  const int* ulen;          // [n_landmarks]
  double* Hrx;              // [slots]
  double* Hrr;              // [n_rho]
  double* yrho;             // [n_rho] back-substituted step
};

// ---- normal-equation assembly by tile gather (assemble.cu)
// one residual table's loss-corrected Jacobian in HBM: one block of `rstride` doubles per residual = [row][column] | r[row] | padding to a
// 16-byte multiple (so a block moves with one bulk async copy); pos [residual][column] = position in the linear system (-1: constant block,
// or a column merged into an earlier one)
// Variable-length sets (the Schur rows of the landmarks): boff[res] = offset of the block / of its positions, bcols[res] = its columns;
// row_only_last: the last column of every row feeds tile ROWS only (the right-hand-side element of a Schur row).
struct RowSetView { const double* J; const int* pos; int ncols, rows, rstride, lo, hi; const int* boff; const int* bcols; int row_only_last; };
LVI_HD int asm_block_doubles(int t) { return (rt_rows(t) * rt_cols(t) + rt_rows(t) + 1) & ~1; }
struct AsmSets { RowSetView s[RT_COUNT]; };
struct AsmPlan {
  DBuf<double> J[RT_COUNT];
  DBuf<int> pos[RT_COUNT];
  DBuf<unsigned long long> keys;   // (tile << 32 | table << 29 | residual), sorted
  DBuf<int> item_start;            // [n_items + 1] offsets into keys: one work item = <= kChunk residuals of one tile
  DBuf<unsigned char> desc;        // [n_entries][64] x desc_width bytes: Jacobian column behind each tile row / tile column, all ones = none
  int desc_width = 1;
  AsmSets sets{};
  int n_items = 0, n_entries = 0;
  bool built = false;
};

// multi-GPU reduction of the normal equations over peer memory (p2p.cu)
struct P2PPlan {
  bool active = false;
  int n_tile_units = 0, n_units = 0, n_list = 0;
  size_t small_elems = 0, off_flags = 0, off_data = 0;
  DBuf<unsigned char> tile_flags;   // [n_tile_units] 1 = this rank's assembly plan writes the tile
  DBuf<unsigned char> tile_flags_all;   // [world][n_tile_units] the same of every rank
  DBuf<int> unit_list;              // units the reduction visits (tiles some rank writes + the small units)
  DBuf<int> own_list;               // units this rank clears before it linearises (its own tiles + the small units)
  int n_own = 0;
  DBuf<double> small_sum;           // reduced {corner, g, Schur rows, Schur diagonal, cost}, copied out to the private buffers
  DBuf<int> counter;
};

// free parameter block (for Plus / norms): kind 0 Euclidean, 1 quaternion (x,y,z,w), 2 Euclidean with lower bound 0
struct FreeBlock { int off; int pos; int size; int kind; };

}  // namespace lvi

struct lvi_problem {
  lvi_ctx* ctx = nullptr;
  lvi_problem_desc desc{};   // caller's arrays (in/out)
  lvi::Lowered L;
  // parameter vectors
  int off_r3 = 0, off_so3 = 0, off_sens = 0, off_rho = 0, nx = 0;
  lvi::DBuf<double> X, XC, XS;        // current, candidate, saved
  lvi::DBuf<lvi::FreeBlock> blocks;   // free blocks
  int n_blocks = 0;
  // tables
  lvi::DBuf<int> tab_i[lvi::RT_COUNT][5];      // i0a, i0b, ia, ib, perm
  lvi::DBuf<double> tab_d[lvi::RT_COUNT][5];   // ua, ub, v, weight, huber
  lvi::DBuf<int> pos_r3, pos_so3, pos_rho;
  lvi::DBuf<double> planes;
  lvi::ProblemView view{};   // device pointers, parameters -> X
  // normal equations
  lvi::BandSys H{}, A{}, A2{};   // A2: second-level system of the separator (two-sided ordering)
  lvi::DBuf<double> H_tiles, H_C, A_tiles, A_C, A_Linv, A_x, A_work_d, A2_tiles, A2_C, A2_Linv, A2_x, A2_work_d;
  lvi::DBuf<int> A_work_i, A2_work_i;
  lvi::DBuf<unsigned long long> A_ll, A2_ll;
  lvi::DBuf<int> pack_map;       // multi-GPU: indices of the structurally non-zero tiles of H (what the all-reduce has to move)
  lvi::DBuf<double> pack_buf;    // ... and their contiguous staging copy
  int n_pack = 0;
  lvi::AsmPlan asmp;
  lvi::AsmPlan schur_plan;             // the same machinery for the Schur complement on the inverse depths: rows = landmarks
  lvi::DBuf<double> schur_rows;        // [blocks] sqrt(1/d) * (merged coupling row | right-hand-side element)
  lvi::DBuf<int> schur_boff, schur_bcols, schur_pos;
  bool schur_gather = false;
  lvi::P2PPlan p2p;
  lvi::SchurView schur{};
  // where a linearisation writes: the private buffers (H, schur, g, scal), or this rank's peer-memory region (multi-GPU, p2p.cu)
  lvi::BandSys H_lin{};
  lvi::SchurView schur_lin{};
  double* g_lin = nullptr;
  double* cost_lin = nullptr;
  lvi::DBuf<int> row_start, row_pos, lm_of_rho, slot2u, urow_pos, ulen;
  lvi::DBuf<double> Hrx, Hrr, yrho;
  lvi::DBuf<int> fail;
  lvi::DBuf<double> g, scale, diag, y, delta, scal;  // scal: small scalar scratch (cost etc.)
  double* h_scal = nullptr;                          // pinned host mirror of scal (the context's)
  int nt = 0;
  bool has_solver_buffers = false;
};

namespace lvi {
// problem.cu
void problem_set_param_source(lvi_problem* p, const double* x_d);                     // point the view at X or XC
void problem_linearize(lvi_problem* p, double* cost_d);                               // zero H,g ; accumulate J^T J, J^T r, cost
void problem_cost(lvi_problem* p, const double* x_d, double* cost_d, bool active_only, bool inactive_only);
void problem_ensure_solver_buffers(lvi_problem* p);
void problem_download_params(lvi_problem* p);
// assemble.cu
void assemble_build_plan(lvi_problem* p);   // once per problem, after problem_ensure_solver_buffers
void assemble_gather(lvi_problem* p);       // H tiles, corner and g from the Jacobian rows jacobian_kernel<TYPE> left in the plan's buffers
void assemble_build_schur_plan(lvi_problem* p);                       // once per problem: landmarks as rows of a second plan (false: rows too long)
void assemble_schur_gather(lvi_problem* p);                           // A -= sum over landmarks of (row)(row)^T from p->schur_rows
void assemble_mark_tiles(lvi_problem* p, unsigned char* flags_d);   // flags[tile] = 1 for every band / border tile this rank's plan writes
// p2p.cu
bool p2p_prepare(lvi_problem* p);           // collective; true: the linearisation targets point into the peer-memory region
void p2p_begin_linearize(lvi_problem* p);
void p2p_reduce(lvi_problem* p);
void p2p_ctx_release(lvi_ctx* ctx);
// solver.cu
void band_factor_solve(lvi_ctx* ctx, BandSys& A, BandSys& A2);
void init_second_level(const BandSys& A, BandSys& A2);   // sizes of the separator system (no allocation)
}  // namespace lvi
