// TEST INFRASTRUCTURE (not product code).
// Force-included into every translation unit of the reference runtime built
// by oracle/Makefile: routes the runtime's console output (VL_PRINTF,
// include/verilated.h:344) into a per-thread capture buffer so the tests can
// compare $display/$write text instead of scraping stdout.
#ifndef VLREF_HOOKS_H_
#define VLREF_HOOKS_H_
#ifdef __cplusplus
extern "C" {
#endif
int vlref_printf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
#ifdef __cplusplus
}
#include <cstdarg>
extern "C" int vlref_vprintf(const char* fmt, va_list ap);
#endif
#define VL_PRINTF vlref_printf
#define VL_VPRINTF vlref_vprintf
#endif
