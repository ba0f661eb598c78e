%% cons_ellipMPC_ADMM_CUDA - constructor of the ellipMPC ADMM  solver for the 'CUDA' platform
% Goes to formulations/+ellipMPC/ ; dispatched by name from spcies_gen_controller.m:114-130.
% Same ingredients and tables as cons_ellipMPC_ADMM_C; kernel template spcies_b200/csrc/MPC_ADMM.cuh.
function constructor = cons_ellipMPC_ADMM_CUDA(recipe)
    constructor = cons_generic_CUDA(recipe, @ellipMPC.cons_ellipMPC_ADMM_C, 'ellipMPC_ADMM', 'MPC_ADMM.cuh', {'#define SPCIES_TERMINAL 2'}, 0);
end
