/* Emits a table of sizeof / offsetof for every ABI struct member in abi_fields.inc.  Compiled twice: once against this
 * repository's include/longtail_abi.h (ABI_TABLE=abi_table_b200) and once against the reference's src/longtail.h
 * (-DLONGTAIL_B200_USE_LONGTAIL_H, ABI_TABLE=abi_table_reference); tests/dropin/dropin_test.c compares the two. */
#include "longtail_abi.h"

#include <stddef.h>

struct abi_entry
{
    const char* name;
    unsigned long value;
};

#define S(s) {"sizeof " #s, (unsigned long)sizeof(struct s)},
#define X(s, m) {#s "." #m, (unsigned long)offsetof(struct s, m)},
const struct abi_entry ABI_TABLE[] = {
#include "abi_fields.inc"
    {0, 0}};
