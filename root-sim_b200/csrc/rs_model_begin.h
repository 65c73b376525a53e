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
/* libc of a model whose INIT parses an input file (models/traffic/init.c): in-memory stdio over the images the
 * host loaded (rs_prelude.cuh), string functions with device bodies */
#define FILE			rs_FILE
#define fopen(p, m)		rs_fopen((p), (m))
#define fclose(f)		rs_fclose(f)
#define fgets(b, n, f)		rs_fgets((b), (n), (f))
#define fgetc(f)		rs_fgetc(f)
#define feof(f)			rs_feof(f)
#define rewind(f)		rs_rewind(f)
#define perror(s)		rs_perror(s)
#define strlen(s)		rs_strlen(s)
#define strcmp(a, b)		rs_strcmp((a), (b))
#define strncmp(a, b, n)	rs_strncmp((a), (b), (n))
#define strcpy(d, s)		rs_strcpy((d), (s))
#define strncpy(d, s, n)	rs_strncpy((d), (s), (n))
#define strtok_r(s, d, p)	rs_strtok_r((s), (d), (p))
#define strtod(s, e)		rs_strtod((s), (e))
#define atoi(s)			rs_atoi(s)
/* C semantics of abs(): the argument is converted to int (models/traffic/functions.c:125 relies on it) */
#define abs(x)			rs_c_abs((int)(x))
/* log/exp steer timestamps and coin tosses: the portable versions give the same bits on the device, in the
 * emulation and in the CPU oracle (rs_math.h) */
#define log(x)			rs_log(x)
#define exp(x)			rs_exp(x)
