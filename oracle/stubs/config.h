/* Hand-written stand-in for the autoconf-generated config.h of the reference
 * (autotools are not installed here).  TEST INFRASTRUCTURE ONLY: used solely to
 * compile the unmodified reference sources into oracle/_ref/.
 * HAVE_SSE is deliberately NOT defined: times_128 and times_sse give identical
 * results (reference unit_tests/test_rectangular_binary_matrix.cc:153). */
#ifndef JF_ORACLE_CONFIG_H
#define JF_ORACLE_CONFIG_H
#define HAVE_INT128 1
#define HAVE_NUMERIC_LIMITS128 1
#define HAVE_POSIX_MEMALIGN 1
#define HAVE_EXECINFO_H 1
#define HAVE_EXT_STDIO_FILEBUF_H 1
#define HAVE_SYS_SYSCALL_H 1
#define HAVE_SYS_PRCTL_H 1
#define HAVE_SI_INT 1
#define PACKAGE_STRING "jellyfish 2.3.1"
#define PACKAGE_VERSION "2.3.1"
#endif
