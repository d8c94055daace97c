// Host check of csrc/std_sort.cuh: the restated libstdc++ std::sort against std::sort itself, on inputs chosen for
// ties (the arrangement of equal keys is the whole point), for sizes around the insertion-sort threshold, and on
// "killer" inputs produced by McIlroy's adversary run against std::sort, which drive the introsort loop to its
// depth limit and into the heap-sort fallback.  Also the round-based schedule the CUDA block uses (all ranges of one
// level of the partition tree, then the next) against the serial one.  No GPU involved.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

static long g_heap_calls = 0;
#define SS_ON_HEAP (++g_heap_calls)
#include "std_sort.cuh"

static unsigned g_rng = 2463534242u;
static unsigned Rnd() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 17; g_rng ^= g_rng << 5; return g_rng; }

static bool Less(ss_word a, ss_word b) { return (a >> 32) < (b >> 32); }

// the schedule of the device: every range of a round, then the ranges they produced
static void SortByRounds(ss_word* v, int n) {
    if (n <= SS_THRESHOLD) { ss_insertion_sort(v, 0, n); return; }
    std::vector<SsRange> cur{SsRange{0, n, 2 * ss_lg(n)}}, next;
    while (!cur.empty()) {
        next.clear();
        for (size_t t = cur.size(); t-- > 0;) {        // backwards: the order inside a round must not matter
            SsRange out[2];
            const int k = ss_step(v, cur[t], out);
            for (int i = 0; i < k; ++i) next.push_back(out[i]);
        }
        cur.swap(next);
    }
}

static int Check(const std::vector<unsigned>& keys, const char* what) {
    const int n = (int)keys.size();
    std::vector<ss_word> a(n), b, c;
    for (int i = 0; i < n; ++i) a[i] = ((ss_word)keys[i] << 32) | (unsigned)i;
    b = a; c = a;
    std::sort(a.begin(), a.end(), Less);
    ss_sort_serial(b.data(), n);
    SortByRounds(c.data(), n);
    int bad = 0;
    for (int i = 0; i < n; ++i) bad += (a[i] != b[i]) + (a[i] != c[i]);
    if (bad) std::printf("%s n %d: %d differing positions\n", what, n, bad);
    return bad != 0;
}

// M. D. McIlroy, "A Killer Adversary for Quicksort": the comparator decides the keys while the sort runs.
static std::vector<int> g_val;
static int g_solid, g_candidate, g_gas;
static bool AdversaryLess(int x, int y) {
    if (g_val[x] == g_gas && g_val[y] == g_gas) { if (x == g_candidate) g_val[x] = g_solid++; else g_val[y] = g_solid++; }
    if (g_val[x] == g_gas) g_candidate = x; else if (g_val[y] == g_gas) g_candidate = y;
    return g_val[x] < g_val[y];
}
static std::vector<unsigned> Killer(int n) {
    std::vector<int> idx(n);
    g_val.assign(n, n - 1);
    g_gas = n - 1; g_solid = 0; g_candidate = 0;
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), AdversaryLess);
    std::vector<unsigned> keys(n);
    for (int i = 0; i < n; ++i) keys[i] = (unsigned)g_val[i];
    return keys;
}

int main() {
    int errors = 0, cases = 0;
    for (int n = 0; n <= 70; ++n)
        for (int distinct : {1, 2, 3, 5, 1000000}) {
            std::vector<unsigned> k(n);
            for (int rep = 0; rep < 20; ++rep) {
                for (int i = 0; i < n; ++i) k[i] = Rnd() % distinct;
                errors += Check(k, "small"); ++cases;
            }
        }
    for (int n : {100, 257, 1000, 4980, 8191, 8192, 65537, 300000})
        for (int distinct : {1, 2, 7, 100, 5000, 1 << 30}) {
            std::vector<unsigned> k(n);
            for (int i = 0; i < n; ++i) k[i] = Rnd() % distinct;
            errors += Check(k, "random"); ++cases;
            std::sort(k.begin(), k.end());
            errors += Check(k, "ascending"); ++cases;
            std::reverse(k.begin(), k.end());
            errors += Check(k, "descending"); ++cases;
            for (int i = 0; i < n; ++i) k[i] = (unsigned)(i < n / 2 ? i : n - i) % distinct;
            errors += Check(k, "organ pipe"); ++cases;
        }
    const long before = g_heap_calls;
    for (int n : {200, 1000, 5000, 40000})
        for (int fold : {1, 2, 5}) {
            std::vector<unsigned> k = Killer(n);
            for (unsigned& x : k) x /= fold;          // fold > 1: the same shape with ties
            errors += Check(k, "killer"); ++cases;
        }
    const long heap = g_heap_calls - before;
    if (heap == 0) { std::printf("the killer inputs never reached the heap-sort fallback\n"); ++errors; }
    // partial_sort(first, last, last) on its own
    for (int n : {2, 3, 16, 17, 100, 1001})
        for (int distinct : {1, 3, 50, 1 << 30}) {
            std::vector<ss_word> a(n), b;
            for (int i = 0; i < n; ++i) a[i] = ((ss_word)(Rnd() % distinct) << 32) | (unsigned)i;
            b = a;
            std::partial_sort(a.begin(), a.end(), a.end(), Less);
            ss_heap_sort(b.data(), n);
            int bad = 0;
            for (int i = 0; i < n; ++i) bad += a[i] != b[i];
            if (bad) std::printf("heap n %d: %d differing positions\n", n, bad);
            errors += bad != 0; ++cases;
        }
    std::printf("%d cases, %ld heap-sort fallbacks, %d errors\n", cases, heap, errors);
    return errors != 0;
}
