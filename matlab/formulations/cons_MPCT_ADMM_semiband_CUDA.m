%% cons_MPCT_ADMM_semiband_CUDA - constructor of the MPCT ADMM semiband solver for the 'CUDA' platform
% Goes to formulations/+MPCT/ ; dispatched by name from spcies_gen_controller.m:114-130.
% Same ingredients and tables as cons_MPCT_ADMM_semiband_C; kernel template spcies_b200/csrc/MPCT_ADMM_semiband.cuh.
function constructor = cons_MPCT_ADMM_semiband_CUDA(recipe)
    constructor = cons_generic_CUDA(recipe, @MPCT.cons_MPCT_ADMM_semiband_C, 'MPCT_ADMM_semiband', 'MPCT_ADMM_semiband.cuh', {}, 0);
end
