// mmpc_modelgen: build-time driver that runs the product's own generator on the built-in models.
//   mmpc_modelgen <model> <out_dir>   writes <out_dir>/mmpc_model_<model>.cuh and <model>_linear_functions.c
#include <cstdio>
#include <fstream>
#include <iostream>

#include "codegen.hpp"

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s <model> <out_dir>\n", argv[0]);
        return 2;
    }
    mahi::mpc::codegen::ModelExpr m;
    if (!mahi::mpc::codegen::builtin_model(argv[1], &m)) {
        std::fprintf(stderr, "unknown built-in model '%s'\n", argv[1]);
        return 1;
    }
    const std::string dir = argv[2];
    std::ofstream(dir + "/mmpc_model_" + m.name + ".cuh") << mahi::mpc::codegen::emit_cuda_model(m);
    std::ofstream(dir + "/" + m.name + "_linear_functions.c") << mahi::mpc::codegen::emit_linear_functions_c(m);
    std::cout << "generated " << m.name << " into " << dir << std::endl;
    return 0;
}
