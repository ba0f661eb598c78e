%% cons_ellipMPC_ADMM_soc_CUDA - constructor of the ellipMPC ADMM soc solver for the 'CUDA' platform
% Goes to formulations/+ellipMPC/ ; dispatched by name from spcies_gen_controller.m:114-130.
% Same ingredients and tables as cons_ellipMPC_ADMM_soc_C; kernel template spcies_b200/csrc/ellipMPC_ADMM_soc.cuh.
function constructor = cons_ellipMPC_ADMM_soc_CUDA(recipe)
    constructor = cons_generic_CUDA(recipe, @ellipMPC.cons_ellipMPC_ADMM_soc_C, 'ellipMPC_ADMM_soc', 'ellipMPC_ADMM_soc.cuh', {}, 1);
end
