%% get_generic_solver_struct - path of the skeleton of a generated .cu file (CUDA platform)
function path_to_file = get_generic_solver_struct()
    full_path = mfilename('fullpath');
    this_path = fileparts(full_path);
    path_to_file = [this_path '/generic_solver_struct.cu'];
end
