// Host-side check of the slot -> pixel maps of csrc/tpt_internal.h (tpt_part_slots / tpt_slot_pixel): for every
// partition kind, rank, world and number of launch chains the slots of all (rank, chain) pairs cover every pixel
// exactly once.  Built with nvcc, runs without a GPU (only __host__ code executes).
#include <cstdio>
#include <cstring>
#include <vector>

#include "tpt_internal.h"

int main() {
    int bad = 0, cases = 0;
    const int sizes[] = {1, 2, 7, 96 * 96, 784 * 784 / 16 + 3};
    for (int npix : sizes)
        for (int partition = TPT_PART_ALL; partition <= TPT_PART_BLOCK; ++partition)
            for (int world = 1; world <= 8; ++world)
                for (int nsub = 1; nsub <= 3; ++nsub) {
                    if (partition == TPT_PART_ALL && world > 1) continue;
                    std::vector<int> seen(npix, 0);
                    for (int rank = 0; rank < world; ++rank) {
                        int total = 0;
                        for (int sub = 0; sub < nsub; ++sub) {
                            RenderArgs a;
                            std::memset(&a, 0, sizeof a);
                            a.partition = partition; a.rank = rank; a.world = world; a.sub = sub; a.nsub = nsub;
                            const int S = tpt_part_slots(a, npix);
                            total += S;
                            for (int slot = 0; slot < S; ++slot) {
                                const int p = tpt_slot_pixel(a, npix, slot);
                                if (p < 0 || p >= npix) { ++bad; continue; }
                                seen[p]++;
                            }
                        }
                        RenderArgs all;
                        std::memset(&all, 0, sizeof all);
                        all.partition = partition; all.rank = rank; all.world = world; all.sub = 0; all.nsub = 1;
                        if (total != tpt_part_slots(all, npix)) ++bad;
                    }
                    for (int p = 0; p < npix; ++p) if (seen[p] != 1) ++bad;
                    ++cases;
                }
    std::printf("%d cases, %d errors\n", cases, bad);
    return bad ? 1 : 0;
}
