// mstep.cuh -- M-step entry points (included at the end of phyfoo.cu).  Placeholder until the
// selection / fixed-q free-energy kernels land: every entry point reports PHYFOO_EUNSUPPORTED.
extern "C" int phyfoo_select(phyfoo_ctx* ctx, const phyfoo_dataset*, int32_t, const double*, double, double, int32_t*, int32_t*,
                             double*, int64_t, int64_t*) { return fail(ctx, PHYFOO_EUNSUPPORTED, "phyfoo_select: not implemented yet"); }
extern "C" int phyfoo_fe_prepare(phyfoo_ctx* ctx, const phyfoo_dataset*, phyfoo_model*, int64_t, const int32_t*, const int32_t*,
                                 const uint8_t*, const double*, phyfoo_feset**) { return fail(ctx, PHYFOO_EUNSUPPORTED, "phyfoo_fe_prepare: not implemented yet"); }
extern "C" int phyfoo_fe_prepare_all(phyfoo_ctx* ctx, const phyfoo_dataset*, phyfoo_model*, const double*, phyfoo_feset**) {
    return fail(ctx, PHYFOO_EUNSUPPORTED, "phyfoo_fe_prepare_all: not implemented yet"); }
extern "C" int phyfoo_fe_eval(phyfoo_ctx* ctx, phyfoo_feset*, int32_t, const double*, int32_t, double*) {
    return fail(ctx, PHYFOO_EUNSUPPORTED, "phyfoo_fe_eval: not implemented yet"); }
extern "C" int phyfoo_fe_grad(phyfoo_ctx* ctx, phyfoo_feset*, int32_t, const double*, int32_t, double, double*, double*) {
    return fail(ctx, PHYFOO_EUNSUPPORTED, "phyfoo_fe_grad: not implemented yet"); }
extern "C" int phyfoo_fe_values_device(phyfoo_ctx* ctx, phyfoo_feset*, int32_t, const double*, int32_t, double, double*, void*) {
    return fail(ctx, PHYFOO_EUNSUPPORTED, "phyfoo_fe_values_device: not implemented yet"); }
extern "C" int64_t phyfoo_fe_size(const phyfoo_feset*) { return 0; }
extern "C" void phyfoo_fe_free(phyfoo_ctx*, phyfoo_feset*) {}
