/* heapstress -- a small ROOT-Sim model written for this repository's tests, after the reference's allocator unit test
 * (tests/dymelor.c: random malloc / calloc / realloc / free over a set of bins, every block filled with a pattern that
 * is checked before the block is touched again).  Here the bins live in the LP's rollbackable heap and the actions
 * are events, so that every rollback has to bring back the allocator's own state (free list, break, block headers)
 * together with the blocks.  The pattern depends on (bin, size, offset) only, never on an address, so that the CPU
 * oracle (libc malloc) and the engine (per-LP arena) produce the same state: `failures` must stay 0 and `sum`, a
 * running hash of everything read back, is compared bit for bit. */
#include <stdlib.h>
#include <string.h>
#include <ROOT-Sim.h>

enum { ACT = 1 };
#define BINS		12
#define MAX_SIZE	600

typedef struct bin {
	unsigned char *ptr;
	unsigned size;
} bin_t;

typedef struct lp_state {
	unsigned ops;			/* events processed                                   byte  0 */
	unsigned failures;		/* pattern or zero checks that failed (must stay 0)         4 */
	unsigned long long sum;		/* running hash of the bytes read back                      8 */
	unsigned live_bytes;		/* bytes in the bins                                        16 */
	unsigned mallocs, callocs, reallocs, frees;	/*                                   20..35 */
	bin_t bins[BINS];
} lp_state_t;

static unsigned char pat(unsigned bin, unsigned size, unsigned i)
{
	unsigned j = (bin * 2654435761u) ^ (size * 40503u) ^ (i * 2246822519u);
	return (unsigned char)(j ^ (j >> 8) ^ (j >> 19));
}

static void mem_init(lp_state_t *s, unsigned b)
{
	unsigned i;
	for (i = 0; i < s->bins[b].size; i++)
		s->bins[b].ptr[i] = pat(b, s->bins[b].size, i);
}

/* check the first n bytes of bin b against the pattern of a block of `size` bytes */
static void mem_check(lp_state_t *s, unsigned b, unsigned size, unsigned n)
{
	unsigned i;
	for (i = 0; i < n; i++) {
		unsigned char c = s->bins[b].ptr[i];
		s->sum = s->sum * 1099511628211ULL + c;
		if (c != pat(b, size, i))
			s->failures++;
	}
}

void ProcessEvent(unsigned int me, simtime_t now, int event_type, void *content, unsigned int size, lp_state_t *s)
{
	simtime_t when;
	unsigned b, act, nsz, old, i, keep;
	unsigned char *p;
	(void)content; (void)size;
	switch (event_type) {
	case INIT:
		s = malloc(sizeof(lp_state_t));
		memset(s, 0, sizeof(lp_state_t));
		SetState(s);
		when = 3 * Random();
		ScheduleNewEvent(me, when, ACT, NULL, 0);
		break;
	case ACT:
		s->ops++;
		b = (unsigned)RandomRange(0, BINS - 1);
		act = (unsigned)RandomRange(0, 3);
		nsz = (unsigned)RandomRange(1, MAX_SIZE);
		if (s->bins[b].ptr == NULL) {
			if (act & 1) {
				p = calloc(nsz, 1);
				s->callocs++;
				for (i = 0; i < nsz; i++)
					if (p[i] != 0)
						s->failures++;
			} else {
				p = malloc(nsz);
				s->mallocs++;
			}
			s->bins[b].ptr = p;
			s->bins[b].size = nsz;
			s->live_bytes += nsz;
			mem_init(s, b);
		} else if (act < 2) {
			mem_check(s, b, s->bins[b].size, s->bins[b].size);
			free(s->bins[b].ptr);
			s->frees++;
			s->live_bytes -= s->bins[b].size;
			s->bins[b].ptr = NULL;
			s->bins[b].size = 0;
		} else {
			old = s->bins[b].size;
			keep = old < nsz ? old : nsz;
			s->bins[b].ptr = realloc(s->bins[b].ptr, nsz);
			s->reallocs++;
			mem_check(s, b, old, keep);		/* realloc keeps the common prefix */
			s->bins[b].size = nsz;
			s->live_bytes += nsz;
			s->live_bytes -= old;
			mem_init(s, b);
		}
		when = now + Expent(1.0);
		if (RandomRange(0, 9) < 3)
			ScheduleNewEvent((unsigned)RandomRange(0, (int)n_prc_tot - 1), when, ACT, NULL, 0);
		else
			ScheduleNewEvent(me, when, ACT, NULL, 0);
		break;
	}
}

bool OnGVT(unsigned int me, lp_state_t *snapshot)
{
	(void)me; (void)snapshot;
	return false;
}
