%% cons_equMPC_FISTA_CUDA - constructor of the equMPC FISTA  solver for the 'CUDA' platform
% Goes to formulations/+equMPC/ ; dispatched by name from spcies_gen_controller.m:114-130.
% Same ingredients and tables as cons_equMPC_FISTA_C; kernel template spcies_b200/csrc/MPC_FISTA.cuh.
function constructor = cons_equMPC_FISTA_CUDA(recipe)
    hdr = 'MPC_FISTA.cuh';
    if recipe.options.time_varying; hdr = 'MPC_FISTA_tv.cuh'; end      % per-instance model: factorisation on the device
    constructor = cons_generic_CUDA(recipe, @equMPC.cons_equMPC_FISTA_C, 'equMPC_FISTA', hdr, {'#define SPCIES_TERMINAL 0'}, 0);
end
