// k_halo.cu -- multi-GPU ghost-body halo (pack / unpack) entry points of the C ABI. See DESIGN.md "Multi-GPU".
#include "b200_context.h"
struct b200phys_context { b200::Context c; };
extern "C" {
int b200phys_halo_pack(b200phys_context*, int, float, float, void**, uint32_t*, uint32_t*) { return B200PHYS_E_INVALID; }
int b200phys_halo_unpack(b200phys_context*, const void*, uint32_t) { return B200PHYS_E_INVALID; }
int b200phys_halo_clear(b200phys_context*) { return B200PHYS_E_INVALID; }
uint32_t b200phys_halo_record_bytes(void) { return 0; }
}
