// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Link shim for building the *unmodified* reference sources with g++/libgomp.
// The reference calls omp_init_lock_with_hint (src/tsxcount/TSXHashMapOMPPerf.h:58),
// an OpenMP 4.5 entry point; some libgomp builds do not export it. A weak
// definition forwards to omp_init_lock when the runtime has none. Nothing else
// of the reference is touched: the sources are compiled where they lie.
#include <omp.h>

extern "C" __attribute__((weak)) void omp_init_lock_with_hint(omp_lock_t* lock, omp_sync_hint_t) {
    omp_init_lock(lock);
}
