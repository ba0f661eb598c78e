%% dec_var - CUDA version
%
% [decl, init] = CUDA_code.dec_var(row)
%   decl: member declaration inside 'struct spcies_consts', e.g. 'SPCIES_REAL Alpha[9][6][6];\n'
%   init: aggregate initialiser of that member,             e.g. '/* Alpha */ {{{R_(0.1..), ...}}},\n'
% Numbers follow platforms/+C_code/dec_var.m: ints '%d', reals '%1.15f', +-inf -> +-1e20; 3-D arrays are
% laid out [dim3][dim1][dim2] (block, row, col). Reals are wrapped in R_() = (SPCIES_REAL)(x) so that
% precision = 'float' rounds decimal -> double -> float exactly like 'float x = 0.123...;' does in C.
%
function [decl, init] = dec_var(var)
    name = var{1}; value = var{2}; type = var{4};
    is_int = any(strcmp(type, {'int', 'uint', 'dint', 'udint', 'sint', 'usint'}));
    if is_int; ctype = 'int'; else; ctype = 'SPCIES_REAL'; end
    dim = size(value);
    if max(dim) == 1
        decl = sprintf('    %s %s;\\n', ctype, name);
        init = sprintf('    /* %s */ %s,\\n', name, wv(value, is_int));
    elseif length(dim) == 2 && min(dim) == 1
        decl = sprintf('    %s %s[%d];\\n', ctype, name, max(dim));
        init = sprintf('    /* %s */ %s,\\n', name, vec(value(:), is_int));
    elseif length(dim) == 2
        decl = sprintf('    %s %s[%d][%d];\\n', ctype, name, dim(1), dim(2));
        init = sprintf('    /* %s */ %s,\\n', name, mat(value, is_int));
    else
        decl = sprintf('    %s %s[%d][%d][%d];\\n', ctype, name, dim(3), dim(1), dim(2));
        blocks = cell(1, dim(3));
        for k = 1:dim(3); blocks{k} = mat(value(:, :, k), is_int); end
        init = sprintf('    /* %s */ {%s},\\n', name, strjoin(blocks, ', '));
    end
end

function s = wv(x, is_int)
    x(x == inf) = 1e20; x(x == -inf) = -1e20;
    if is_int; s = sprintf('%d', x); else; s = sprintf('R_(%1.15f)', x); end
end
function s = vec(v, is_int)
    c = arrayfun(@(x) wv(x, is_int), v, 'UniformOutput', false);
    s = ['{' strjoin(c(:)', ', ') '}'];
end
function s = mat(M, is_int)
    rows = cell(1, size(M, 1));
    for i = 1:size(M, 1); rows{i} = vec(M(i, :), is_int); end
    s = ['{' strjoin(rows, ', ') '}'];
end
