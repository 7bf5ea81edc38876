/*
 * TEST INFRASTRUCTURE ONLY.  Random differential test of the function-select boundary, the GPU analogue of the reference's own
 * hevcasm_test loop over instruction sets (reference hevcasm_test.c:110-137): the per-block cases of the library self-test - every
 * reference partition size, every fractional position, several quantiser settings, all transform sizes - run on FRESH random inputs
 * (one set per seed) through
 *     the reference's own compiled C path, slots populated with HEVCASM_C_REF | HEVCASM_C_OPT   (oracle/_ref/libhevcasm_cref.so)
 *     this repo's library, slots populated with HEVCASM_CUDA                                     (hevcasm_b200/libhevcasm_b200.so)
 * and the 64-bit digests of all outputs must agree slot by slot.  Both libraries export the same symbol names, so each is opened
 * with dlopen(RTLD_LOCAL) and the case code is compiled once per side (ref_harness_side.c).
 *
 *   ref_harness <libhevcasm_cref.so> <libhevcasm_b200.so> [n_seeds = 8] [first_seed = 1]        exit code = number of differing digests
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

int ref_init(void *dl), cuda_init(void *dl), ref_total(void), cuda_total(void);
int ref_run(uint64_t salt, int mask, uint64_t *out), cuda_run(uint64_t salt, int mask, uint64_t *out);

#define HEVCASM_C_REF_OPT 3 /* HEVCASM_C_REF | HEVCASM_C_OPT: X-macro values 0 and 1 (reference hevcasm.h:113-125) */
#define HEVCASM_CUDA_BIT (1 << 9)

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: ref_harness <libhevcasm_cref.so> <libhevcasm_b200.so> [n_seeds] [first_seed]\n");
        return 2;
    }
    const int n_seeds = argc > 3 ? atoi(argv[3]) : 8;
    const uint64_t first = argc > 4 ? (uint64_t)atoll(argv[4]) : 1;
    void *dl_ref = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL), *dl_cuda = dlopen(argv[2], RTLD_NOW | RTLD_LOCAL);
    if (!dl_ref || !dl_cuda) {
        fprintf(stderr, "ref_harness: dlopen failed: %s\n", dlerror());
        return 2;
    }
    if (ref_init(dl_ref) || cuda_init(dl_cuda)) return 2;
    const int total = ref_total();
    uint64_t *a = (uint64_t *)calloc((size_t)total, sizeof *a), *b = (uint64_t *)calloc((size_t)total, sizeof *b);
    int bad = 0;
    for (int s = 0; s < n_seeds; ++s) {
        const uint64_t salt = first + (uint64_t)s;
        const int ma = ref_run(salt, HEVCASM_C_REF_OPT, a), mb = cuda_run(salt, HEVCASM_CUDA_BIT, b);
        if (ma || mb) {
            fprintf(stderr, "seed %llu: %d reference slots and %d CUDA slots are empty\n", (unsigned long long)salt, ma, mb);
            bad += ma + mb;
        }
        int diff = 0;
        for (int i = 0; i < total; ++i)
            if (a[i] != b[i]) {
                if (++diff <= 8) fprintf(stderr, "seed %llu case %d: reference %016llx, CUDA %016llx\n", (unsigned long long)salt, i, (unsigned long long)a[i], (unsigned long long)b[i]);
            }
        bad += diff;
        printf("seed %llu: %d cases, %d differ\n", (unsigned long long)salt, total, diff);
    }
    printf("ref_harness: %d seeds x %d cases through C_REF|C_OPT and HEVCASM_CUDA slots: %d differences\n", n_seeds, total, bad);
    return bad > 255 ? 255 : bad;
}
