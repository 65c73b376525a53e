/* rngtail -- a small ROOT-Sim model written for this repository's tests: every event draws from the WHOLE
 * numerical library of the model API (src/lib/numerical.c:80-249): Random, RandomRange, RandomRangeNonUniform,
 * Expent, Normal, Gamma (both branches), Poisson, Zipf.  Timestamps and receivers are derived only from the
 * draws that are bit-exact between the CPU and the device (integer core, log, sqrt); the draws that go through
 * exp()/pow() (Gamma(ia >= 6), Zipf) are kept in the state, where the tests compare them with a stated ULP
 * tolerance, and their integer results and the RNG stream position they leave behind are compared exactly. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <ROOT-Sim.h>

enum { STEP = 1 };

typedef struct lp_state {
	unsigned n;			/* events processed                                    byte  0 */
	int rr;				/* last RandomRange(0, 99)                                   4 */
	int rrnu;			/* last RandomRangeNonUniform(7, 3, 40)                      8 */
	int zipf;			/* last Zipf(1.5, 1000)                                     12 */
	unsigned long long isum;	/* running mix of all integer draws                         16 */
	double normal;			/* last Normal()          (log, sqrt: bit-exact)            24 */
	double gamma_small;		/* last Gamma(3)          (log: bit-exact)                  32 */
	double poisson;			/* last Poisson()         (log: bit-exact)                  40 */
	double gamma_big;		/* last Gamma(8)          (exp: ULP tolerance)              48 */
} lp_state_t;

unsigned step_limit = 200;

void ProcessEvent(unsigned int me, simtime_t now, int event_type, void *content, unsigned int size, lp_state_t *s)
{
	simtime_t when;
	double g, gs, e;
	int to;
	(void)content; (void)size;
	switch (event_type) {
	case INIT:
		s = malloc(sizeof(lp_state_t));
		memset(s, 0, sizeof(lp_state_t));
		SetState(s);
		when = 5 * Random();
		ScheduleNewEvent(me, when, STEP, NULL, 0);
		break;
	case STEP:
		s->n++;
		s->rr = RandomRange(0, 99);
		s->rrnu = RandomRangeNonUniform(7, 3, 40);
		g = Normal();
		gs = Gamma(3);
		s->normal = g;
		s->gamma_small = gs;
		s->poisson = Poisson();
		s->gamma_big = Gamma(8);
		s->zipf = Zipf(1.5, 1000);
		s->isum = s->isum * 1000003ULL + (unsigned long long)(s->rr * 10007 + s->rrnu * 101 + s->zipf);
		e = Expent(2.0);
		when = now + 0.25 + e + fabs(g) * 0.5 + gs * 0.125;
		to = RandomRange(0, (int)n_prc_tot - 1);
		if (s->rr < 30)
			ScheduleNewEvent((unsigned)to, when, STEP, NULL, 0);
		else
			ScheduleNewEvent(me, when, STEP, NULL, 0);
		break;
	}
}

bool OnGVT(unsigned int me, lp_state_t *snapshot)
{
	(void)me;
	return snapshot->n >= step_limit;
}
