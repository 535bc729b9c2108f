/*
 * ntransformer.h — public C API of the engine.
 *
 * Same declarations as the reference's include/ntransformer.h:11-38, which the reference declares but never
 * defines; libnt_b200.so implements them over the resident B200 engine.  Strings returned by
 * nt_engine_generate are malloc'ed and must be released with nt_free.
 */
#ifndef NTRANSFORMER_H
#define NTRANSFORMER_H

#ifdef __cplusplus
extern "C" {
#endif

#include <stddef.h>
#include <stdint.h>

typedef void* nt_engine_t;

nt_engine_t nt_engine_create(void);
void        nt_engine_destroy(nt_engine_t engine);

/* returns 0 on success */
int nt_engine_load(nt_engine_t engine, const char* model_path);

char* nt_engine_generate(nt_engine_t engine, const char* prompt, int max_tokens,
                         float temperature, int top_k, float top_p);
void nt_free(char* ptr);

int nt_engine_vocab_size(nt_engine_t engine);
int nt_engine_n_layers(nt_engine_t engine);
int nt_engine_hidden_size(nt_engine_t engine);

#ifdef __cplusplus
}
#endif
#endif /* NTRANSFORMER_H */
