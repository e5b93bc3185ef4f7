// Variable-length peptide table + tryptic digest/lookup kernel (prot2tryp2lca.rs:88-140).
#include "index.h"

namespace umgap {
int build_var_table_from_fst(const char* path, umgap_index* idx, double load_factor) {
    (void)path; (void)idx; (void)load_factor;
    set_error("variable-length (tryptic) table not built yet");
    return UMGAP_ERR_INVALID;
}
}  // namespace umgap

extern "C" {
uint64_t umgap_tryp_lookup_bound(uint64_t total_aa, uint64_t nlines) { return total_aa + nlines + 1; }
int umgap_tryp_lookup(const umgap_index*, const uint8_t*, const uint64_t*, uint64_t, int, int, const char*,
                      const char*, int, uint32_t*, uint64_t*) {
    umgap::set_error("umgap_tryp_lookup not implemented yet");
    return UMGAP_ERR_INVALID;
}
}
