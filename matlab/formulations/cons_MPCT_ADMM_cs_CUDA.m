%% cons_MPCT_ADMM_cs_CUDA - constructor of the MPCT ADMM cs solver for the 'CUDA' platform
% Goes to formulations/+MPCT/ ; dispatched by name from spcies_gen_controller.m:114-130.
% Same ingredients and tables as cons_MPCT_ADMM_cs_C; kernel template spcies_b200/csrc/MPCT_ADMM_cs.cuh.
function constructor = cons_MPCT_ADMM_cs_CUDA(recipe)
    constructor = cons_generic_CUDA(recipe, @MPCT.cons_MPCT_ADMM_cs_C, 'MPCT_ADMM_cs', 'MPCT_ADMM_cs.cuh', {}, 0);
end
