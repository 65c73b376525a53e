/* oracle/kat_ref_stubs.c -- TEST INFRASTRUCTURE ONLY: link-time stubs for oracle/kat_ref_driver.c */
#include <stdarg.h>
#include <stdbool.h>
__thread unsigned int __lp_counter; unsigned int n_prc;
void *lps_blocks;
char rootsim_config[4096];
void _mkdir(const char *p) { (void)p; }
void _rootsim_error(bool fatal, const char *msg, ...) { (void)fatal; (void)msg; }
