// vbk_symbolic.cpp -- see vbk_symbolic.h.  Host-side, runs once per LP.
#include "vbk_symbolic.h"

#include <algorithm>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>

namespace vbk {

namespace {
constexpr long long kSymCacheMagic = 0x324d59534b4256ll;   // "VBKSYM2": ordering only

// Indexed binary min-heap, positions 1..count.  Tie behaviour must equal the reference's static
// hfall/hrise pair (ldlt.c:1305-1349): the right child wins only when strictly smaller, and an
// element moves only when strictly out of order.  Any other choice changes perm[].
class KeyHeap {
public:
    explicit KeyHeap(int n) : key(n), where(n), at(n + 1), count(n) {}
    std::vector<int> key;    // key per node
    std::vector<int> where;  // heap position of node
    std::vector<int> at;     // node at heap position (1-based)
    int count;

    int top() const { return at[1]; }

    void sift_down(int pos) {
        for (int child = 2 * pos; child <= count; child = 2 * pos) {
            if (child < count && key[at[child + 1]] < key[at[child]]) ++child;
            if (key[at[pos]] > key[at[child]]) { exchange(pos, child); pos = child; }
            else break;
        }
    }
    void sift_up(int pos) {
        for (int up = pos / 2; up > 0; up = pos / 2) {
            if (key[at[up]] > key[at[pos]]) { exchange(pos, up); pos = up; }
            else break;
        }
    }
    // ldlt.c:1125-1133: overwrite the hole with the last element, shrink, then restore order
    // (sink if the removed key was smaller than the moved one, otherwise float).
    void remove(int node) {
        int pos = where[node];
        int removed_key = key[at[pos]];
        at[pos] = at[count];
        where[at[pos]] = pos;
        --count;
        if (removed_key < key[at[pos]]) sift_down(pos);
        else sift_up(pos);
    }

private:
    void exchange(int a, int b) {
        std::swap(at[a], at[b]);
        where[at[a]] = a;
        where[at[b]] = b;
    }
};

}  // namespace

// Tiered minimum-degree ordering with mass elimination on an explicit-fill elimination graph.
// Follows ldlt.c:860-1262 decision for decision (method = minimum degree).
void Symbolic::order(std::vector<std::vector<int>>& adj, std::vector<int>& tier) {
    const int penalty = (int)(1.0 * N);  // stablty * m, ldlt.c:889
    perm.assign(N, -1);
    iperm.assign(N, -1);
    kL.assign(N + 1, 0);
    iL.clear();
    {
        long long half = 0;
        for (auto& a : adj) half += (long long)a.size();
        iL.reserve((size_t)half);  // grows as fill is discovered
    }
    std::vector<int> stamp(N, 0), estamp(N, 0), others;
    others.reserve(N);

    KeyHeap heap(N);
    for (int v = 0; v < N; ++v) heap.key[v] = (int)adj[v].size();
    for (int v = 0; v < N; ++v) {
        if ((int)adj[v].size() > dense && tier[v] == 0) tier[v] = 1;  // ldlt.c:994-999
        heap.key[v] += tier[v] * penalty;
    }
    for (int v = N - 1; v >= 0; --v) {  // ldlt.c:1004-1010
        heap.where[v] = v + 1;
        heap.at[v + 1] = v;
        heap.sift_down(v + 1);
    }

    // Eliminated nodes are NOT removed from their neighbours' lists on the spot (the reference shifts them out,
    // ldlt.c:1094-1120): they are recognised by gone[] wherever a list is read and squeezed out the next time the list is
    // scanned for fill; deg[] carries the live degrees.  clique[u] = the elimination step whose survivors u belonged to
    // last: that step made its survivors pairwise adjacent, and edges between live nodes never disappear, so a later
    // step whose survivors all carry one common clique id has no fill to find and no list to scan -- once the graph has
    // become dense (the heavy top of the elimination tree) a step costs O(degree) instead of O(degree^2).
    std::vector<int> deg(N), clique(N, -1);
    std::vector<char> gone(N, 0);
    for (int v = 0; v < N; ++v) deg[v] = (int)adj[v].size();
    int tag = 0, etag = 0;
    denwin = N;
    double tph[6] = {0,0,0,0,0,0};
    auto now = []{ return std::chrono::steady_clock::now(); };
    const bool timing = std::getenv("VBK_SYM_STATS") != nullptr;
    for (int i = 0; i < N;) {
        auto tp0 = now();
        const int pivot = heap.top();
        const int d = deg[pivot];
        if (d >= N - 1 - i) denwin = i;  // ldlt.c:1027
        perm[i] = pivot;
        iperm[pivot] = i;
        {   // the pivot's own list is read several times below: squeeze the eliminated nodes out of it first
            auto& lst = adj[pivot];
            size_t w = 0;
            for (size_t k = 0; k < lst.size(); ++k) if (!gone[lst[k]]) lst[w++] = lst[k];
            lst.resize(w);
        }

        // neighbours indistinguishable from the pivot join its group (ldlt.c:1037-1054)
        for (int u : adj[pivot]) iperm[u] = i;
        others.clear();
        int iend = i + 1;
        for (int u : adj[pivot]) {
            bool same = (deg[u] == d) && tier[u] == tier[pivot];
            if (same) {
                for (int w : adj[u]) if (!gone[w] && iperm[w] < i) { same = false; break; }
            }
            if (same) { perm[iend] = u; iperm[u] = iend; ++iend; }
            else others.push_back(u);
        }

        auto tp1 = now();
        // column structures of the group members (ldlt.c:1068-1088); old node ids for now
        for (int ii = i, len = d; ii < iend; ++ii, --len) {
            kL[ii + 1] = kL[ii] + len;
            for (int w : adj[perm[ii]]) {
                if (gone[w]) continue;
                int row = iperm[w];
                if (row > ii || (row == i && w != pivot)) iL.push_back(w);
            }
        }

        auto tp2 = now();
        for (int ii = i; ii < iend; ++ii) heap.remove(perm[ii]);  // ldlt.c:1122-1134
        auto tp3 = now();

        // every group member is adjacent to every survivor (its neighbourhood is the pivot's): they leave the graph
        const int nelim = iend - i;
        for (int ii = i; ii < iend; ++ii) gone[perm[ii]] = 1;
        const int nsurv = (int)others.size();
        for (int u : others) deg[u] -= nelim;

        // pairwise fill among the survivors, appended at the list ends (ldlt.c:1144-1201)
        bool one_clique = nsurv > 0 && clique[others[0]] >= 0;
        for (int a = 1; a < nsurv && one_clique; ++a) one_clique = clique[others[a]] == clique[others[0]];
        if (!one_clique) {
            ++etag;
            for (int u : others) estamp[u] = etag;
            for (int a = 0; a < nsurv; ++a) {
                const int u = others[a];
                auto& lst = adj[u];
                ++tag;
                int have = 0;
                size_t w = 0;
                for (size_t k = 0; k < lst.size(); ++k) {
                    const int v = lst[k];
                    if (gone[v]) continue;                // eliminated, in this step or an earlier one
                    lst[w++] = v;
                    if (stamp[v] != tag) { stamp[v] = tag; have += (estamp[v] == etag); }    // a repeated entry counts once
                }
                lst.resize(w);
                if (have == nsurv - 1) continue;          // u already sees every other survivor
                for (int b = a + 1; b < nsurv; ++b) {
                    const int x = others[b];
                    if (stamp[x] != tag) { lst.push_back(x); adj[x].push_back(u); ++deg[u]; ++deg[x]; }
                }
            }
            for (int u : others) clique[u] = i;
        }

        auto tp4 = now();
        // re-key survivors in list order: float, then sink (ldlt.c:1206-1220)
        for (int u : others) {
            heap.key[u] = deg[u] + (tier[u] != 0 ? tier[u] * penalty : 0);
            heap.sift_up(heap.where[u]);
            heap.sift_down(heap.where[u]);
        }

        for (int ii = i; ii < iend; ++ii) std::vector<int>().swap(adj[perm[ii]]);
        i = iend;
        if (timing) { auto tp5 = now(); auto D = [](auto a, auto b){ return std::chrono::duration<double>(b - a).count(); };
            tph[0] += D(tp0, tp1); tph[1] += D(tp1, tp2); tph[2] += D(tp2, tp3); tph[3] += D(tp3, tp4); tph[4] += D(tp4, tp5); }
    }

    if (timing) std::fprintf(stderr, "vbk ordering phases: group test %.3f, column structures %.3f, heap remove %.3f, fill %.3f, re-key %.3f s\n", tph[0], tph[1], tph[2], tph[3], tph[4]);
    auto tq0 = now();
    for (int& r : iL) r = iperm[r];  // ldlt.c:1236
    {
        // rows ascending inside each column (ldlt.c:1238 sorts every column; the result only depends on the sets): two
        // bucket passes -- by row, then back by column in ascending row order -- instead of a sort per column
        const size_t nnz = iL.size();
        std::vector<int> rcnt((size_t)N + 1, 0), rcol(nnz);
        for (size_t k = 0; k < nnz; ++k) rcnt[(size_t)iL[k] + 1]++;
        for (int r = 0; r < N; ++r) rcnt[r + 1] += rcnt[r];
        {
            std::vector<int> fill(rcnt.begin(), rcnt.end() - 1);
            for (int j = 0; j < N; ++j)
                for (int k = kL[j]; k < kL[j + 1]; ++k) rcol[(size_t)fill[iL[k]]++] = j;
        }
        std::vector<int> fill(kL.begin(), kL.end() - 1);
        for (int r = 0; r < N; ++r)
            for (int t = rcnt[r]; t < rcnt[r + 1]; ++t) iL[(size_t)fill[rcol[t]]++] = r;
    }
    iL.shrink_to_fit();
    if (timing) std::fprintf(stderr, "vbk ordering: map + sort columns %.3f s\n", std::chrono::duration<double>(now() - tq0).count());

    narth = 0.0;
    for (int j = 0; j < N; ++j) { double c = kL[j + 1] - kL[j]; narth += c * c; }
    narth = narth + 3.0 * kL[N] + N;
}

// Fill pattern of L from the ordering alone (used when the ordering comes from the disk cache): column j of L holds the
// rows > j of K's permuted column j merged with the patterns of j's elimination-tree children -- the classic symbolic
// factorisation, O(Lnz).  The reference's explicit-fill elimination (ldlt.c:1050-1236) produces exactly this set, rows
// ascending (checked against it on the netlib fixtures, tests/test_host.py).
void Symbolic::pattern_from_ordering(const std::vector<std::vector<int>>& adj, size_t lnz_hint) {
    std::vector<std::vector<int>> kids(N);
    std::vector<int> mark(N, -1);
    kL.assign((size_t)N + 1, 0);
    iL.clear();
    iL.reserve(lnz_hint);
    for (int j = 0; j < N; ++j) {
        const size_t b0 = iL.size();
        mark[j] = j;
        for (int u : adj[perm[j]]) {
            const int r = iperm[u];
            if (r > j && mark[r] != j) { mark[r] = j; iL.push_back(r); }
        }
        for (int k : kids[j]) {
            for (int q = kL[k]; q < kL[k + 1]; ++q) {
                const int r = iL[q];
                if (mark[r] != j) { mark[r] = j; iL.push_back(r); }
            }
        }
        std::vector<int>().swap(kids[j]);
        std::sort(iL.begin() + b0, iL.end());
        if (iL.size() > 0x7fffffffu) { std::fprintf(stderr, "vbkkt: L has more than 2^31-1 entries (32-bit ABI, SURVEY H5)\n"); std::exit(1); }
        kL[j + 1] = (int)iL.size();
        if (iL.size() > b0) kids[iL[b0]].push_back(j);
    }
}

void Symbolic::analyze(int m_, int n_, const int* kA, const int* iA, const int* kAt, const int* iAt) {
    m = m_; n = n_; N = m + n; nzA = kA[n];

    // ordering priority from the two fill estimates (ldlt.c:687-717); double arithmetic in the
    // reference's loop order because a borderline comparison could flip otherwise
    double fraction = 1.0e0;
    for (int j = 0; j < n; ++j) {
        double dens = (double)(kA[j + 1] - kA[j]) / (m + 1);
        fraction = fraction * (1.0e0 - dens * dens);
    }
    const double pfillin = 0.5 * m * m * (1.0e0 - fraction);
    fraction = 1.0e0;
    for (int i = 0; i < m; ++i) {
        double dens = (double)(kAt[i + 1] - kAt[i]) / (n + 1);
        fraction = fraction * (1.0e0 - dens * dens);
    }
    const double dfillin = 0.5 * n * n * (1.0e0 - fraction);
    pdf = (3 * pfillin <= dfillin) ? 1 : 2;  // Q is empty, so "separable" is true

    // adjacency of K in the reference's initial neighbour order (ldlt.c:727-759)
    std::vector<std::vector<int>> adj(N);
    for (int j = 0; j < n; ++j) {
        adj[j].reserve(2 * (size_t)(kA[j + 1] - kA[j]));
        for (int k = kA[j]; k < kA[j + 1]; ++k) adj[j].push_back(n + iA[k]);
    }
    for (int i = 0; i < m; ++i) {
        adj[n + i].reserve(2 * (size_t)(kAt[i + 1] - kAt[i]));
        adj[n + i].assign(iAt + kAt[i], iAt + kAt[i + 1]);
    }
    // tiers (ldlt.c:766-809) for bndmark=BDD_BELOW, rngmark=INFINITE: favoured side 0, other 1
    std::vector<int> tier(N);
    for (int j = 0; j < n; ++j) tier[j] = (pdf == 1) ? 0 : 1;
    for (int i = 0; i < m; ++i) tier[n + i] = (pdf == 1) ? 1 : 0;
    dense = 3;  // ldlt.c:814-846 with n1 == 0

    const auto t0 = std::chrono::steady_clock::now();
    // $VBK_SYM_CACHE=<directory>: the ordering (perm, iperm, denwin, narth: 8 bytes per row/column of K) is kept on disk,
    // keyed by a hash of the matrix pattern; the fill pattern is rebuilt from it by an elimination-tree pass in O(Lnz) (the explicit-fill ordering of the largest synthetic LPs
    // takes minutes to an hour, for the reference as for any faithful restatement, and depends on the pattern only)
    std::string cache_file;
    if (const char* dir = std::getenv("VBK_SYM_CACHE")) {
        unsigned long long h = 1469598103934665603ull;
        auto mix = [&h](const void* p, size_t bytes) {
            const unsigned char* c = static_cast<const unsigned char*>(p);
            for (size_t i = 0; i < bytes; ++i) { h ^= c[i]; h *= 1099511628211ull; }
        };
        mix(&m, sizeof m); mix(&n, sizeof n); mix(kA, sizeof(int) * ((size_t)n + 1)); mix(iA, sizeof(int) * (size_t)nzA);
        char name[64];
        std::snprintf(name, sizeof name, "/vbksym_%d_%d_%016llx.bin", m, n, h);
        cache_file = std::string(dir) + name;
    }
    bool cached = false;
    if (!cache_file.empty()) {
        if (FILE* f = std::fopen(cache_file.c_str(), "rb")) {
            long long hd[8] = {0};
            double na = 0;
            if (std::fread(hd, sizeof hd, 1, f) == 1 && std::fread(&na, sizeof na, 1, f) == 1 && hd[0] == kSymCacheMagic &&
                hd[1] == m && hd[2] == n && hd[3] == nzA && hd[4] == N) {
                perm.resize(N); iperm.resize(N);
                cached = std::fread(perm.data(), sizeof(int), N, f) == (size_t)N && std::fread(iperm.data(), sizeof(int), N, f) == (size_t)N;
                if (cached) {
                    pattern_from_ordering(adj, (size_t)hd[5]);          // the fill pattern is a function of the ordering
                    cached = (long long)kL[N] == hd[5];
                }
                denwin = (int)hd[6]; narth = na;
                if (!cached) std::fprintf(stderr, "vbkkt: symbolic cache %s is damaged; recomputing\n", cache_file.c_str());
            }
            std::fclose(f);
        }
    }
    if (!cached) {
        order(adj, tier);
        if (!cache_file.empty()) {
            const std::string tmp = cache_file + ".tmp";
            if (FILE* f = std::fopen(tmp.c_str(), "wb")) {
                const long long hd[8] = {kSymCacheMagic, m, n, nzA, N, (long long)kL[N], denwin, pdf};
                bool ok = std::fwrite(hd, sizeof hd, 1, f) == 1 && std::fwrite(&narth, sizeof narth, 1, f) == 1 &&
                          std::fwrite(perm.data(), sizeof(int), N, f) == (size_t)N && std::fwrite(iperm.data(), sizeof(int), N, f) == (size_t)N;
                ok = (std::fclose(f) == 0) && ok;
                if (ok) std::rename(tmp.c_str(), cache_file.c_str()); else std::remove(tmp.c_str());
            }
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    derive(kA, iA, kAt, iAt);
    if (std::getenv("VBK_SYM_STATS"))
        std::fprintf(stderr, "vbk symbolic: ordering %.3f s, derived structures %.3f s (N %d, Lnz %d)\n",
                     std::chrono::duration<double>(t1 - t0).count(),
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count(), N, kL[N]);
}

void Symbolic::derive(const int* kA, const int* iA, const int* kAt, const int* iAt) {
    const int lnz_ = kL[N];

    // elimination tree: rows ascend inside a column, so the parent is the first stored row
    parent.assign(N, -1);
    nchild.assign(N, 0);
    height.assign(N, 0);
    maxcol = 0;
    for (int j = 0; j < N; ++j) {
        int c = kL[j + 1] - kL[j];
        maxcol = std::max(maxcol, c);
        if (c > 0) { parent[j] = iL[kL[j]]; nchild[parent[j]]++; }
    }
    nlevels = 0;
    for (int j = 0; j < N; ++j) {  // children precede parents
        if (parent[j] >= 0) height[parent[j]] = std::max(height[parent[j]], height[j] + 1);
        nlevels = std::max(nlevels, height[j] + 1);
    }
    lvlptr.assign(nlevels + 1, 0);
    for (int j = 0; j < N; ++j) lvlptr[height[j] + 1]++;
    for (int l = 0; l < nlevels; ++l) lvlptr[l + 1] += lvlptr[l];
    lvlcol.assign(N, 0);
    {
        std::vector<int> fill(lvlptr.begin(), lvlptr.end() - 1);
        for (int j = 0; j < N; ++j) lvlcol[fill[height[j]]++] = j;
    }

    // row lists of L
    rowptr.assign(N + 1, 0);
    for (int k = 0; k < lnz_; ++k) rowptr[iL[k] + 1]++;
    for (int r = 0; r < N; ++r) rowptr[r + 1] += rowptr[r];
    rk_asc.assign(lnz_, 0); rj_asc.assign(lnz_, 0);
    rk_sig.assign(lnz_, 0); rj_sig.assign(lnz_, 0);
    {
        std::vector<int> fill(rowptr.begin(), rowptr.end() - 1);
        for (int j = 0; j < N; ++j)
            for (int k = kL[j]; k < kL[j + 1]; ++k) {
                int dst = fill[iL[k]]++;
                rk_asc[dst] = k;
                rj_asc[dst] = j;
            }
    }
    {
        // Integer skeleton of lltnum's first/link bookkeeping (ldlt.c:550-552,568-580,616-620):
        // the order in which row r meets its columns is the order every sum of column r is
        // accumulated in, so the numeric kernels replay exactly this list.
        std::vector<int> cursor(N, 0), chain(N, -1);
        for (int r = 0; r < N; ++r) {
            int dst = rowptr[r];
            for (int j = chain[r], nextj; j != -1; j = nextj) {
                nextj = chain[j];
                int k = cursor[j];
                rk_sig[dst] = k;
                rj_sig[dst] = j;
                ++dst;
                if (k + 1 < kL[j + 1]) {
                    cursor[j] = k + 1;
                    int row = iL[k + 1];
                    chain[j] = chain[row];
                    chain[row] = j;
                }
            }
            if (dst != rowptr[r + 1]) { std::fprintf(stderr, "vbk: row list mismatch at %d\n", r); std::abort(); }
            if (kL[r] < kL[r + 1]) {
                cursor[r] = kL[r];
                int row = iL[kL[r]];
                chain[r] = chain[row];
                chain[row] = r;
            }
        }
    }

    // scatter maps (ldlt.c:243-269): K entry (row>col) lands at the L slot of that (row, col)
    mapA.assign(nzA, -1);
    mapAt.assign(nzA, -1);
    {
        std::vector<int> slot(N, -1);
        for (int j = 0; j < n; ++j) {
            int col = iperm[j];
            for (int k = kL[col]; k < kL[col + 1]; ++k) slot[iL[k]] = k;
            for (int k = kA[j]; k < kA[j + 1]; ++k) {
                int row = iperm[n + iA[k]];
                if (row > col) mapA[k] = slot[row];
            }
        }
        for (int i = 0; i < m; ++i) {
            int col = iperm[n + i];
            for (int k = kL[col]; k < kL[col + 1]; ++k) slot[iL[k]] = k;
            for (int k = kAt[i]; k < kAt[i + 1]; ++k) {
                int row = iperm[iAt[k]];
                if (row > col) mapAt[k] = slot[row];
            }
        }
    }

    // task decomposition for the numeric kernels (see vbk_symbolic.h)
    if (const char* e = std::getenv("VBK_WHOLE_CAP")) whole_cap = std::max(1, std::atoi(e)); else whole_cap = 32;
    slice_row0 = N;
    for (int j = 0; j < N; ++j)
        if (kL[j + 1] - kL[j] > whole_cap) { slice_row0 = j; break; }
    if (const char* e = std::getenv("VBK_ROWBLK")) rowblk = std::max(1, std::atoi(e)); else rowblk = 32;   // measured on B200: 32 beats 64 (profiles/)
    nblk = 0;
    tasks_ok = true;
    winptr.clear();
    if (slice_row0 < N) {
        nblk = (N - slice_row0 + rowblk - 1) / rowblk;
        while ((long long)N * (nblk + 1) > (1LL << 28) && rowblk < 128) {   // keep the lookup table below 1 GiB
            rowblk *= 2;
            nblk = (N - slice_row0 + rowblk - 1) / rowblk;
        }
        if ((long long)N * (nblk + 1) > (1LL << 28)) {
            // The dense per-column block table does not fit at the kernel's largest row block (128): no slice tasks for
            // this LP.  Fast mode factorises its sparse columns with the level kernels; strict mode refuses the LP.
            tasks_ok = false;
            nblk = 0;
        }
    }
    if (slice_row0 < N && tasks_ok) {
        winptr.assign((size_t)N * (nblk + 1), 0);
        for (int j = 0; j < N; ++j) {
            int* wp = &winptr[(size_t)j * (nblk + 1)];
            int k = kL[j];
            for (int b = 0; b <= nblk; ++b) {
                long long bound = (long long)slice_row0 + (long long)b * rowblk;
                while (k < kL[j + 1] && iL[k] < bound) ++k;
                wp[b] = (b == nblk) ? kL[j + 1] : k;
            }
        }
    }
    task_col.clear(); task_blk.clear(); task_pos0.clear(); task_cnt.clear();
    col_task0.assign(N, 0);
    col_ntask.assign(N, 0);
    for (int j = 0; j < N && tasks_ok; ++j) {
        col_task0[j] = (int)task_col.size();
        const int c = kL[j + 1] - kL[j];
        if (c <= whole_cap) {
            task_col.push_back(j); task_blk.push_back(-1); task_pos0.push_back(kL[j]); task_cnt.push_back(c);
        } else {
            const int* wp = &winptr[(size_t)j * (nblk + 1)];
            for (int b = 0; b < nblk; ++b) {
                int p0 = wp[b], p1 = wp[b + 1];
                if (p1 > p0) {
                    task_col.push_back(j); task_blk.push_back(b); task_pos0.push_back(p0); task_cnt.push_back(p1 - p0);
                }
            }
        }
        col_ntask[j] = (int)task_col.size() - col_task0[j];
    }

    if (std::getenv("VBK_SYM_STATS") && std::atoi(std::getenv("VBK_SYM_STATS")) >= 2) {
        // how much of the factorisation's work lies in contributor segments that cover every row of their task
        long long full_pairs = 0, part_pairs = 0, empty_pairs = 0, full_prod = 0, part_prod = 0, run_prod = 0;
        for (int t = 0; t < (int)task_col.size(); ++t) {
            const int i = task_col[t], blk = task_blk[t], c = task_cnt[t];
            for (int q = rowptr[i]; q < rowptr[i + 1]; ++q) {
                const int j = rj_sig[q];
                int kb = rk_sig[q] + 1, ke = kL[j + 1];
                if (blk >= 0) {
                    const int* wp = &winptr[(size_t)j * (nblk + 1)];
                    kb = std::max(kb, wp[blk]); ke = std::min(ke, wp[blk + 1]);
                }
                const int len = std::max(0, ke - kb);
                if (len == 0) ++empty_pairs;
                else if (len == c) { ++full_pairs; full_prod += len; }
                else {
                    ++part_pairs; part_prod += len;
                    if (iL[ke - 1] - iL[kb] == len - 1) run_prod += len;      // contiguous rows
                }
            }
        }
        std::fprintf(stderr, "vbk symbolic stats: tasks %zu, (task, contributor) pairs: full %lld, partial %lld, empty %lld; products: full %lld, partial %lld (of which contiguous runs %lld)\n",
                     task_col.size(), full_pairs, part_pairs, empty_pairs, full_prod, part_prod, run_prod);
    }

    // Trailing window treated densely by fast mode.  rho = 1 gives the reference's dense window
    // (every column full); smaller rho admits columns that hold at least rho of the rows below them.
    // Padding is exact: the fill pattern is closed under elimination, so entries outside it receive
    // no contribution and stay 0.  The window is limited by memory (W^2 doubles <= 4 GiB) and by
    // work (dense flops W^3/3 at most 8x the factorisation's own flop count).
    {
        // default: 0.06 for factorisations that are worth it (>= 1e8 flops: dfl001 7.0 vs 9.4 ms per KKT step, pilot87
        // 3.7 vs 6.1, multicommodity R=32 39 vs 44 -- the sparse tree loses its deep, heavy top, profiles/r01_summary.md),
        // 0.25 for small ones, where the window is not what costs time and a narrower tolerance part keeps more of the
        // reference's arithmetic (the fast-mode sweep over netlib was validated with it)
        double rho = narth >= 1.0e8 ? 0.06 : 0.25;
        if (const char* e = std::getenv("VBK_WINDOW_RHO")) rho = std::atof(e);
        for (;;) {
            dense_start = N;
            for (int j = N - 1; j >= 0; --j) {
                if ((double)(kL[j + 1] - kL[j]) >= rho * (double)(N - 1 - j)) dense_start = j; else break;
            }
            const double W = (double)(N - dense_start);
            if (rho >= 1.0 || (W * W * 8.0 <= 4.0 * 1073741824.0 && W * W * W / 3.0 <= 8.0 * narth + 1.0e6)) break;
            rho = std::min(1.0, rho * 2.0);
        }
    }

    // fundamental supernodes: column j+1 continues j's supernode when j+1 is j's only-child parent
    // and struct(j+1) = struct(j) \ {j+1}
    sn_ptr.clear();
    sn_of.assign(N, 0);
    for (int j = 0; j < N; ++j) {
        bool cont = false;
        if (j > 0) {
            int cprev = kL[j] - kL[j - 1], ccur = kL[j + 1] - kL[j];
            cont = parent[j - 1] == j && nchild[j] == 1 && ccur == cprev - 1;
        }
        if (!cont) sn_ptr.push_back(j);
        sn_of[j] = (int)sn_ptr.size() - 1;
    }
    sn_ptr.push_back(N);
}

}  // namespace vbk
