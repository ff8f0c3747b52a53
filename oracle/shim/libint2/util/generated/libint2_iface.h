/* oracle/shim/libint2/util/generated/libint2_iface.h -- TEST INFRASTRUCTURE (oracle).
 *
 * Stand-in for the generator-emitted interface header; declarations follow
 * /root/reference/src/bin/libint/iface.cc:114-185 (tables, static init/cleanup,
 * per-task init/need_memory/cleanup) and :257-286 (the LIBINT2_PREFIXED_NAME /
 * LIBINT2_DEFINED macros). Definitions live in oracle/oracle_kernels.cc.
 */
#ifndef _libint2_oracle_iface_h_
#define _libint2_oracle_iface_h_
#include <libint2/util/generated/libint2_types.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
extern void (*libint2_build_default[LIBINT2_MAX_AM_default + 1][LIBINT2_MAX_AM_default + 1])(const Libint_t*);
extern void (*libint2_build_eri[LIBINT2_MAX_AM_eri + 1][LIBINT2_MAX_AM_eri + 1][LIBINT2_MAX_AM_eri + 1][LIBINT2_MAX_AM_eri + 1])(const Libint_t*);
extern void (*libint2_build_3eri[LIBINT2_MAX_AM_3eri + 1][LIBINT2_MAX_AM_3eri + 1][LIBINT2_MAX_AM_3eri + 1])(const Libint_t*);
extern void (*libint2_build_2eri[LIBINT2_MAX_AM_2eri + 1][LIBINT2_MAX_AM_2eri + 1])(const Libint_t*);
void libint2_static_init();
void libint2_static_cleanup();
void libint2_init_default(Libint_t* inteval, int max_am, void* buf);
size_t libint2_need_memory_default(int max_am);
void libint2_cleanup_default(Libint_t* inteval);
void libint2_init_eri(Libint_t* inteval, int max_am, void* buf);
size_t libint2_need_memory_eri(int max_am);
void libint2_cleanup_eri(Libint_t* inteval);
void libint2_init_3eri(Libint_t* inteval, int max_am, void* buf);
size_t libint2_need_memory_3eri(int max_am);
void libint2_cleanup_3eri(Libint_t* inteval);
void libint2_init_2eri(Libint_t* inteval, int max_am, void* buf);
size_t libint2_need_memory_2eri(int max_am);
void libint2_cleanup_2eri(Libint_t* inteval);
#ifdef __cplusplus
}
#endif
#define LIBINT2_PREFIXED_NAME(name) __libint2_prefixed_name__(LIBINT2_API_PREFIX, name)
#define __libint2_prefixed_name__(prefix, name) __prescanned_prefixed_name__(prefix, name)
#define __prescanned_prefixed_name__(prefix, name) prefix##name
#define LIBINT2_DEFINED(taskname, symbol) __prescanned_libint2_defined__(taskname, symbol)
#define __prescanned_libint2_defined__(taskname, symbol) LIBINT2_DEFINED_##symbol
#endif
