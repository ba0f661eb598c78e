%% cons_generic_CUDA - constructor shared by every cons_<F>_<method>[_<sub>]_CUDA
%
% constructor = cons_generic_CUDA(recipe, cons_C, func_name, kernel_hdr, switches, has_r)
%
% Reuses the C platform's constructor of the same solver to obtain the recipe tables (so the ingredients and
% the '#define' / constant rows are, by construction, the ones the plain-C solver gets), then swaps the file
% skeletons: the header keeps the reference's defines and sol_<name> struct and gains the batched prototype
% (SPCIES_CUDA_DECLARE_SOLVER of include/spcies_cuda.h); the code file becomes generic_solver_struct.cu, which
% holds the constants as 'struct spcies_consts' and #includes the hand-written kernel template.
%
% INPUTS:
%   - recipe:     Spcies_problem
%   - cons_C:     handle of the C constructor, e.g. @laxMPC.cons_laxMPC_FISTA_C
%   - func_name:  reference function name, e.g. 'laxMPC_FISTA'
%   - kernel_hdr: kernel template under spcies_b200/csrc, e.g. 'MPC_FISTA.cuh'
%   - switches:   cellstr of extra '#define's for the kernel, e.g. {'#define SPCIES_TERMINAL 1'}
%   - has_r:      true for solvers with the extra r_ellip input (ellipMPC_ADMM_soc)
%
function constructor = cons_generic_CUDA(recipe, cons_C, func_name, kernel_hdr, switches, has_r)
    cuda_root = getenv('SPCIES_CUDA_ROOT');
    rc = recipe.copy(); rc.options.platform = 'C';
    constructor = cons_C(rc);                         % tables: $INSERT_DEFINES$, $INSERT_CONSTANTS$[, $INSERT_VARIABLES$]
    % --- code file: CUDA skeleton
    constructor.files.code.dir.extension = 'cu';
    constructor.files.code.blocks = {'$START$', CUDA_code.get_generic_solver_struct};
    key = recipe.options.formulation;
    if ~isempty(recipe.options.method); key = [key '_' recipe.options.method]; end
    if ~isempty(recipe.options.submethod); key = [key '_' recipe.options.submethod]; end
    constructor.files.code.flags = {'$INSERT_PRECISION$', recipe.options.precision; ...
                                    '$INSERT_FUNC$', func_name; ...
                                    '$INSERT_SOLVER_KEY$', key; ...
                                    '$INSERT_HAS_R$', num2str(has_r); ...
                                    '$INSERT_SWITCHES$', strjoin(switches, '\\n'); ...
                                    '$INSERT_KERNEL$', kernel_hdr};
    % --- constants and variables: members + initialisers of struct spcies_consts
    rows = [];
    for j = 1:size(constructor.data, 1)
        if ~strcmp(constructor.data{j, 1}, '$INSERT_DEFINES$')
            rows = [rows; constructor.data{j, 2}]; %#ok<AGROW>
        end
    end
    text = CUDA_code.declare_variables(rows);
    split = find(strcmp(text, '$SPLIT$'));
    constructor.files.code.flags(end+1, :) = {'$INSERT_MEMBERS$', strjoin(text(1:split-1), '')};
    constructor.files.code.flags(end+1, :) = {'$INSERT_INITIALISERS$', strjoin(text(split+1:end), '')};
    constructor.data = constructor.data(strcmp(constructor.data(:, 1), '$INSERT_DEFINES$'), :);
    % --- header: keep the reference header (defines + sol struct), replace the prototypes by the C-ABI macro
    if has_r; macro = 'SPCIES_CUDA_DECLARE_SOLVER_R'; else; macro = 'SPCIES_CUDA_DECLARE_SOLVER'; end
    constructor.files.header.appends = {};
    constructor.files.header.flags = {'#endif', sprintf(['#include "spcies_cuda.h"\\n#ifdef __cplusplus\\nextern "C" {\\n#endif\\n' ...
        '%s(%s, sol_$INSERT_NAME$);\\n#ifdef __cplusplus\\n}\\n#endif\\n#endif'], macro, func_name)};
    % --- build
    constructor.files.code.exec_me = CUDA_code.get_nvcc_exec(cuda_root);
end
