/*
 * hevcasm_b200 - root header of the B200-native (CUDA sm_100a) build of the HEVCasm kernel library.
 *
 * Declaration-compatible with the reference's src/lib/hevcasm.h (cited per item below) so that a C or
 * C++ codec written against the reference's function-select API compiles and links unchanged; written
 * afresh for this project.  The one extension is a tenth instruction-set entry, HEVCASM_CUDA: slots
 * populated under that bit dispatch to the GPU.  There is no CPU implementation in this library - with
 * any other mask every slot is populated with a null pointer ("no implementation", the reference's own
 * convention, e.g. reference sad.c:207).
 *
 * Batched, stream-aware entry points (the real GPU workload) are in hevcasm_batch.h.
 */
#ifndef INCLUDED_hevcasm_h
#define INCLUDED_hevcasm_h

#include <stddef.h>
#include <stdint.h>
#include <inttypes.h>
#include <stdio.h>

/* reference hevcasm.h:61-69, :104-105 - export / alignment decorations */
#if defined(_WIN32) && defined(HEVCASM_DLL_EXPORTS)
#  define HEVCASM_API __declspec(dllexport)
#elif defined(_WIN32) && defined(HEVCASM_DLL_IMPORTS)
#  define HEVCASM_API __declspec(dllimport)
#else
#  define HEVCASM_API
#endif

#if defined(_MSC_VER)
#  define HEVCASM_ALIGN(n, T, v) __declspec(align(n)) T v
#else
#  define HEVCASM_ALIGN(n, T, v) T v __attribute__((aligned(n)))
#endif

#if defined(__x86_64__) || defined(_WIN64) || defined(__aarch64__)
#  define HEVCASM_X64
#endif

/* reference hevcasm.h:55-58, :86-100 exposes an rdtsc timestamp; here it is a monotonic nanosecond clock
 * (GPU work is timed with CUDA events, see hevcasm_batch.h). */
typedef uint64_t hevcasm_timestamp;

/*
 * Instruction sets as an X-macro: X(bit, NAME, description) - reference hevcasm.h:113-124.
 * Bits 0-8 keep the reference's meaning and are never implemented by this library; bit 9 is the GPU.
 */
#define HEVCASM_INSTRUCTION_SET_XMACRO \
    X(0, C_REF, "C - reference; may be slow") \
    X(1, C_OPT, "C - optimised") \
    X(2, SSE2, "SSE2") \
    X(3, SSE3, "SSE3") \
    X(4, SSSE3, "Supplementary SSE3") \
    X(5, SSE41, "SSE4.1") \
    X(6, SSE42, "SSE4.2") \
    X(7, AVX, "AVX") \
    X(8, AVX2, "AVX2") \
    X(9, CUDA, "NVIDIA CUDA sm_100a (B200)")

#define HEVCASM_INSTRUCTION_SET_COUNT 10

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
#define X(value, name, description) HEVCASM_##name = 1 << value,
    HEVCASM_INSTRUCTION_SET_XMACRO
#undef X
} hevcasm_instruction_set;

hevcasm_timestamp HEVCASM_API hevcasm_get_timestamp_ns(void);

/* reference hevcasm.h:144 / hevcasm.c:99-138 (cpuid probe).  Here: HEVCASM_CUDA when a compute-capability
 * 10.x device is visible, otherwise 0 - this build has nothing to offer a machine without one. */
hevcasm_instruction_set HEVCASM_API hevcasm_instruction_set_support(void);

/* reference hevcasm.h:147 / hevcasm.c:141-149 */
void HEVCASM_API hevcasm_print_instruction_set_support(FILE *f, hevcasm_instruction_set mask);

/* reference hevcasm.h:153 / hevcasm.c:152-186: library self-test; returns the number of errors */
int HEVCASM_API hevcasm_main(int argc, const char *argv[]);

/* reference hevcasm.h:156 */
#define HEVCASM_RECT(width, height) (((width) << 8) | (height))

/* reference hevcasm.h:159 */
typedef void HEVCASM_API hevcasm_test_function(int *error_count, hevcasm_instruction_set mask);

#ifdef __cplusplus
}
#endif

#endif
