%% cons_ellipHMPC_ADMM_CUDA - constructor of the ellipHMPC ADMM solver for the 'CUDA' platform
% Goes to formulations/+HMPC/ ; dispatched by name from spcies_gen_controller.m:114-130.
% Same ingredients and tables as cons_ellipHMPC_ADMM_C; kernel template spcies_b200/csrc/HMPC_ADMM.cuh.
function constructor = cons_ellipHMPC_ADMM_CUDA(recipe)
    constructor = cons_generic_CUDA(recipe, @HMPC.cons_ellipHMPC_ADMM_C, 'ellipHMPC_ADMM', 'HMPC_ADMM.cuh', {'#define SPCIES_NREF 3'}, 0);
end
