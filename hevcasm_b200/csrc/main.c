/* hevcasm_b200 - self-test executable, the counterpart of the reference's src/bin/main.c:39-42.
 * Plain C99 on purpose: it is also the proof that the public headers under include/ are consumable from C. */
#include "hevcasm.h"

int main(int argc, const char *argv[])
{
    return hevcasm_main(argc, argv);
}
