/*
 * rs_model_begin.h -- included by the generated unity source right before the model's own files.
 * Redirects the libc entry points a ROOT-Sim model may call from inside an event to their
 * rollbackable / device-safe counterparts.  This is the compile-time equivalent of the link-time
 * `ld --wrap malloc --wrap free --wrap realloc --wrap calloc` of the reference's rootsim-cc
 * (scripts/rootsim-cc.in:262 -> src/mm/dymelor.c:566-695) and of its str/mem wrappers (:261).
 * Everything defined here is undone by rs_model_end.h.
 */
#define malloc(sz)		(rs_anyptr{rs_malloc(sz)})
#define calloc(n, sz)		(rs_anyptr{rs_calloc((n), (sz))})
#define realloc(p, sz)		(rs_anyptr{rs_realloc((p), (sz))})
#define free(p)			rs_free(p)
#define exit(code)		rs_model_exit(code)
#define abort()			rs_model_exit(-1)
#define fflush(f)		((void)0)
#define memcmp(a, b, n)		rs_memcmp((a), (b), (n))
#define bzero(p, n)		rs_bzero((p), (n))
#define fprintf(stream, ...)	printf(__VA_ARGS__)
/* x86 inline assembly (e.g. models/phold/functions_app.c set_mem) has no device meaning: the
 * statement is dropped, the function keeps its C fallthrough behaviour. */
#define __asm__
#define __volatile__(...)
