/* oracle/shims/numa.h -- single-node stand-in for libnuma (TEST INFRASTRUCTURE ONLY).
 * The reference sizes its OpenMP teams from numa_num_configured_cpus()
 * (core/FullyRepGraph.hpp:36, core/ntsFastSampler.hpp:105); NTS_ORACLE_CPUS overrides. */
#ifndef NTS_ORACLE_SHIM_NUMA_H
#define NTS_ORACLE_SHIM_NUMA_H
#include <stdlib.h>
#include <unistd.h>
struct bitmask { unsigned long size; unsigned long *maskp; };
static inline int numa_available(void) { return 0; }
static inline int numa_num_configured_cpus(void) {
  const char *e = getenv("NTS_ORACLE_CPUS");
  if (e && atoi(e) > 0) return atoi(e);
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
static inline int numa_num_configured_nodes(void) { return 1; }
static inline void *numa_alloc_onnode(size_t sz, int) { return calloc(1, sz ? sz : 1); }
static inline void *numa_alloc_interleaved(size_t sz) { return calloc(1, sz ? sz : 1); }
static inline void *numa_realloc(void *p, size_t, size_t n) { return realloc(p, n); }
static inline void numa_free(void *p, size_t) { free(p); }
static inline void numa_tonode_memory(void *, size_t, int) {}
static inline int numa_run_on_node(int) { return 0; }
static inline struct bitmask *numa_parse_nodestring(const char *) { static struct bitmask b = {0, 0}; return &b; }
static inline void numa_set_interleave_mask(struct bitmask *) {}
#endif
