%% cons_laxMPC_ADMM_CUDA - constructor of the laxMPC ADMM  solver for the 'CUDA' platform
% Goes to formulations/+laxMPC/ ; dispatched by name from spcies_gen_controller.m:114-130.
% Same ingredients and tables as cons_laxMPC_ADMM_C; kernel template spcies_b200/csrc/MPC_ADMM.cuh.
function constructor = cons_laxMPC_ADMM_CUDA(recipe)
    hdr = 'MPC_ADMM.cuh';
    if recipe.options.time_varying; hdr = 'MPC_ADMM_tv.cuh'; end      % per-instance model: factorisation on the device
    constructor = cons_generic_CUDA(recipe, @laxMPC.cons_laxMPC_ADMM_C, 'laxMPC_ADMM', hdr, {'#define SPCIES_TERMINAL 1'}, 0);
end
