/* svm.h -- drop-in replacement for the slice of libsvm 3.17's C API that mipgen.cpp uses
 * (reference svm.h:12-16, 74-98; call sites mipgen.cpp:409, 1951, 2001-2016).
 *
 *   svm_load_model  parses the text model and uploads it to the GPU (mg_load_svr_model)
 *   svm_predict     returns the RBF-SVR decision value computed by K-svr on the GPU
 *
 * struct svm_model is opaque here: mipgen.cpp only ever holds the pointer. */
#ifndef MIPGEN_B200_DROPIN_SVM_H
#define MIPGEN_B200_DROPIN_SVM_H

#define LIBSVM_VERSION 317

#ifdef __cplusplus
extern "C" {
#endif

struct svm_node
{
    int index;    /* 1-based feature index, -1 terminates the vector */
    double value;
};

struct svm_model;

struct svm_model *svm_load_model(const char *model_file_name); /* NULL on failure */
double svm_predict(const struct svm_model *model, const struct svm_node *x);
void svm_free_and_destroy_model(struct svm_model **model_ptr_ptr);
int svm_get_nr_sv(const struct svm_model *model);

#ifdef __cplusplus
}
#endif
#endif
