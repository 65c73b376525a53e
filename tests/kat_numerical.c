/* Known-answer driver for the oracle's restatement of src/lib/numerical.c (TEST INFRASTRUCTURE).
 * Includes oracle/seqsim.c itself (its main renamed) so that the very functions the oracle runs are the
 * ones called here, with the RNG state set directly to the seed of the reference's tests/numerical.c:149.
 * The expected values were produced by the reference's own compiled numerical.o. */
#define main seqsim_main
#include "../oracle/seqsim.c"
#undef main

void ProcessEvent_light(unsigned int me, simtime_t now, int event_type, void *event_content, unsigned int size, void *state)
{
	(void)me; (void)now; (void)event_type; (void)event_content; (void)size; (void)state;
}
bool OnGVT_light(unsigned int me, void *snapshot) { (void)me; (void)snapshot; return true; }

int main(void)
{
	static lp_t lp;
	memset(&lp, 0, sizeof lp);
	lp.seed = 7319936632422683443ULL;
	cur = &lp;
	for (int i = 0; i < 4; i++) {
		double r = Random();
		printf("random %.17g %llu\n", r, (unsigned long long)lp.seed);
	}
	double e = Expent(5.0);
	printf("expent %.17g %llu\n", e, (unsigned long long)lp.seed);
	printf("randomrange %d\n", RandomRange(0, 99));
	double n1 = Normal();
	double n2 = Normal();
	printf("normal %.17g %.17g\n", n1, n2);
	/* the tail: self-consistency values recorded from this restatement (libm flavour), pinned so that a
	 * change of the port shows up */
	printf("rrnu %d\n", RandomRangeNonUniform(7, 3, 40));
	printf("gamma3 %.17g\n", Gamma(3));
	printf("gamma8 %.17g\n", Gamma(8));
	printf("poisson %.17g\n", Poisson());
	printf("zipf %d\n", Zipf(1.5, 1000));
	printf("seed %llu\n", (unsigned long long)lp.seed);
	for (unsigned g = 0; g < 3; g++)
		printf("lpseed %u %llu\n", g, (unsigned long long)lp_seed(12345678901234567ULL, g));
	return 0;
}
