/* check_rng_scale.c — TEST INFRASTRUCTURE.  Exhaustive proof that the device's
 *     (float)((double)x * (1.0 / 4294967295.0))
 * equals the reference's GetRandomFloat (global.cpp:19-22)
 *     (float)((double)x / 0xffffffff)
 * for every 32-bit x.  Build: gcc -O2 -fopenmp check_rng_scale.c -o check_rng_scale */
#include <stdint.h>
#include <stdio.h>
int main(void) {
    const double inv = 2.3283064370807974e-10;
    unsigned long long bad = 0;
    if (inv != 1.0 / 4294967295.0) { printf("constant is not the nearest double\n"); return 2; }
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (long long i = 0; i <= 0xffffffffLL; ++i) {
        uint32_t x = (uint32_t)i;
        volatile float a = (float)((double)x / 0xffffffff);
        volatile float b = (float)((double)x * inv);
        if (a != b) bad++;
    }
    printf("mismatches over 2^32 inputs: %llu\n", bad);
    return bad != 0;
}
