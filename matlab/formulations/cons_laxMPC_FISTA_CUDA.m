%% cons_laxMPC_FISTA_CUDA - constructor of the laxMPC FISTA  solver for the 'CUDA' platform
% Goes to formulations/+laxMPC/ ; dispatched by name from spcies_gen_controller.m:114-130.
% Same ingredients and tables as cons_laxMPC_FISTA_C; kernel template spcies_b200/csrc/MPC_FISTA.cuh.
function constructor = cons_laxMPC_FISTA_CUDA(recipe)
    hdr = 'MPC_FISTA.cuh';
    if recipe.options.time_varying; hdr = 'MPC_FISTA_tv.cuh'; end      % per-instance model: factorisation on the device
    constructor = cons_generic_CUDA(recipe, @laxMPC.cons_laxMPC_FISTA_C, 'laxMPC_FISTA', hdr, {'#define SPCIES_TERMINAL 1'}, 0);
end
