/* oracle/faketime_shim.c -- TEST INFRASTRUCTURE.
 * LD_PRELOAD shim: the SNAP node2vec binary seeds both of its TRnd generators
 * from time(NULL) (bin@0x40c63a and inside LearnEmbeddings); interposing
 * time() makes the unmodified reference binary a deterministic oracle. */
#include <time.h>
#include <stdlib.h>
time_t time(time_t *t) {
    const char *s = getenv("N2V_FAKE_TIME");
    time_t v = s ? (time_t)atol(s) : (time_t)1;
    if (t) *t = v;
    return v;
}
