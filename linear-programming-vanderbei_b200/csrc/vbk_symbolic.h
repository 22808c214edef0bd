// vbk_symbolic.h -- one-time symbolic analysis of the quasidefinite KKT matrix (host, C++).
//
// Product code (not the oracle).  Reproduces, bit for bit, the integer outputs of the reference's
// inv_sym + lltsym (reference src/ipo/ldlt.c:638-1262): perm, iperm, kAAt, iAAt, denwin, pdf.
// On top of those it derives what the GPU numeric phase needs and the reference never builds:
// elimination tree, per-row contribution lists in the reference's accumulation order (SURVEY.md
// section 10), ascending row lists for the triangular solves, the A -> L scatter maps, etree levels
// and the fundamental supernode partition.
#pragma once
#include <cstdint>
#include <vector>

namespace vbk {

struct Symbolic {
    // ldlt-space dimensions (reference ldlt.c:124-136): n "column" nodes 0..n-1 carry -dn,
    // m "row" nodes n..n+m-1 carry +dm.  hsd.c:218 calls with (n_solver, m_solver) swapped.
    int m = 0, n = 0, N = 0;
    int nzA = 0;
    int pdf = 0;      // 1 primal, 2 dual ordering priority (ldlt.c:708-716)
    int dense = 3;    // ldlt.c:814-846 always yields 3 on the ipo path
    int denwin = 0;   // ldlt.c:1027
    double narth = 0; // ldlt.c:1243-1248, the reference's own flop count per factorisation

    std::vector<int> perm, iperm;   // new = iperm[old]
    std::vector<int> kL, iL;        // kAAt[N+1], iAAt[Lnz]; rows ascending inside each column
    int lnz() const { return kL.empty() ? 0 : kL[N]; }

    // ---- derived structures (not in the reference) ----
    std::vector<int> parent;        // etree: first sub-diagonal row of column j, -1 for roots
    std::vector<int> nchild;        // number of etree children
    std::vector<int> height;        // 0 for leaves, 1+max(children) otherwise
    int nlevels = 0;
    std::vector<int> lvlptr, lvlcol; // columns grouped by height (ascending), CSR style
    int maxcol = 0;                 // longest column of L

    // row r of L: entries (column j, position k in the L value array)
    std::vector<int> rowptr;        // [N+1]
    std::vector<int> rk_sig, rj_sig; // in the order lltnum's link lists visit them (ldlt.c:568-580)
    std::vector<int> rk_asc, rj_asc; // ascending column order (rawsolve forward sweep order)

    // scatter maps for inv_num (ldlt.c:243-269): L position of every stored entry of A / At, or -1
    std::vector<int> mapA, mapAt;

    // ---- task decomposition of the numeric factorisation (GPU scheduling, not in the reference) ----
    // A column with at most whole_cap rows is one task.  Longer columns are cut by global row
    // blocks of `rowblk` rows counted from row `slice_row0` (the first long column): one task per
    // block in which the column has rows, so that several CTAs share a long column.  winptr[j*(nblk+1)+b]
    // is the first position in column j whose row is >= slice_row0 + b*rowblk: every contributing
    // column finds its entries for a block by lookup, never by search.
    int whole_cap = 512;              // $VBK_WHOLE_CAP overrides (tests force slicing on tiny LPs)
    int slice_row0 = 0, rowblk = 64, nblk = 0;
    bool tasks_ok = true;             // false: the block table would not fit (huge LPs) -- no slice tasks, fast mode only
    std::vector<int> winptr;          // [N*(nblk+1)], empty when no column is sliced
    std::vector<int> task_col, task_blk, task_pos0, task_cnt;   // blk = -1: whole column
    std::vector<int> col_task0, col_ntask;                      // [N]
    int ntasks() const { return (int)task_col.size(); }
    // first column of the trailing dense window: every column j >= dense_start holds all rows j+1..N-1
    // (the reference's denwin, ldlt.c:1027, re-derived from the pattern so that it can be trusted)
    int dense_start = 0;

    // fundamental supernodes: column ranges [sn_ptr[s], sn_ptr[s+1])
    std::vector<int> sn_ptr;
    std::vector<int> sn_of;         // supernode of each column

    // Runs the analysis.  kA/iA: CSC of the ldlt-space "A" (m rows, n columns); kAt/iAt its transpose.
    void analyze(int m, int n, const int* kA, const int* iA, const int* kAt, const int* iAt);

private:
    void order(std::vector<std::vector<int>>& adj, std::vector<int>& tier);
    void pattern_from_ordering(const std::vector<std::vector<int>>& adj, size_t lnz_hint);
    void derive(const int* kA, const int* iA, const int* kAt, const int* iAt);
};

}  // namespace vbk
