/* rs_model_end.h -- closes the scope opened by rs_model_begin.h. */
#undef malloc
#undef calloc
#undef realloc
#undef free
#undef exit
#undef abort
#undef fflush
#undef memcmp
#undef bzero
#undef fprintf
#undef __asm__
#undef __volatile__
