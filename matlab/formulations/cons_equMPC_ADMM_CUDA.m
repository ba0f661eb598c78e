%% cons_equMPC_ADMM_CUDA - constructor of the equMPC ADMM  solver for the 'CUDA' platform
% Goes to formulations/+equMPC/ ; dispatched by name from spcies_gen_controller.m:114-130.
% Same ingredients and tables as cons_equMPC_ADMM_C; kernel template spcies_b200/csrc/MPC_ADMM.cuh.
function constructor = cons_equMPC_ADMM_CUDA(recipe)
    hdr = 'MPC_ADMM.cuh';
    if recipe.options.time_varying; hdr = 'MPC_ADMM_tv.cuh'; end      % per-instance model: factorisation on the device
    constructor = cons_generic_CUDA(recipe, @equMPC.cons_equMPC_ADMM_C, 'equMPC_ADMM', hdr, {'#define SPCIES_TERMINAL 0'}, 0);
end
