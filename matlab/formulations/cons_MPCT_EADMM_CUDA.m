%% cons_MPCT_EADMM_CUDA - constructor of the MPCT EADMM  solver for the 'CUDA' platform
% Goes to formulations/+MPCT/ ; dispatched by name from spcies_gen_controller.m:114-130.
% Same ingredients and tables as cons_MPCT_EADMM_C; kernel template spcies_b200/csrc/MPCT_EADMM.cuh.
function constructor = cons_MPCT_EADMM_CUDA(recipe)
    constructor = cons_generic_CUDA(recipe, @MPCT.cons_MPCT_EADMM_C, 'MPCT_EADMM', 'MPCT_EADMM.cuh', {}, 0);
end
