%% declare_variables - CUDA version
%
% Declares the variables of a recipe table for the 'CUDA' platform.
% '#define' rows are emitted exactly like the C platform does (same names, same %1.15f text), so the
% generated header carries the reference's defines. Every other row becomes (i) a member declaration of
% 'struct spcies_consts' and (ii) one aggregate initialiser of 'spcies_h_consts'; the two lists are
% returned interleaved as {decl_1; ...; decl_n; '$SPLIT$'; init_1; ...; init_n} and split by
% CUDA_code.cons_generic.
%
% INPUTS:
%   - vars: cell array, one row per variable: {name, value, initialize, type, options}
%           (see platforms/+C_code/dec_var.m for the meaning of the columns)
% OUTPUTS:
%   - s: cell array of strings
%
function s = declare_variables(vars)
    numVars = size(vars, 1);
    decl = {}; init = {}; defs = {};
    for i = 1:numVars
        row = vars(i, :);
        opts = '';
        if length(row) > 4; opts = row{5}; end
        if any(strcmp(opts, 'define'))
            defs{end+1} = C_code.dec_var(row); %#ok<AGROW>  identical text to the C platform
        else
            [d, v] = CUDA_code.dec_var(row);
            decl{end+1} = d; init{end+1} = v; %#ok<AGROW>
        end
    end
    if isempty(decl)
        s = defs;
    else
        s = [defs, decl, {'$SPLIT$'}, init];
    end
end
