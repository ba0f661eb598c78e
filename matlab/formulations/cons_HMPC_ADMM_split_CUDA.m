%% cons_HMPC_ADMM_split_CUDA - constructor of the HMPC ADMM split solver for the 'CUDA' platform
% Goes to formulations/+HMPC/ ; dispatched by name from spcies_gen_controller.m:114-130.
% Same ingredients and tables as cons_HMPC_ADMM_split_C; kernel template spcies_b200/csrc/HMPC_ADMM_split.cuh.
function constructor = cons_HMPC_ADMM_split_CUDA(recipe)
    constructor = cons_generic_CUDA(recipe, @HMPC.cons_HMPC_ADMM_split_C, 'HMPC_ADMM', 'HMPC_ADMM_split.cuh', {}, 0);
end
