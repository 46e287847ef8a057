#!/usr/bin/env python3
"""Turn a scratch copy of the reference tree into the CUDA-enabled TensorFrost module.

The reference has no backend registry: backends and kernel languages are enums plus `switch`
statements (Backend/Backend.h:22-35, Backend/Backend.cpp:49-68,78-89, CodeGen/Generators.cpp:10-22)
and the pybind module binds both enums by value (Frontend/Python/PybindModule.cpp:52-58,76-83).
Adding a backend therefore means adding enum values and switch cases.  This script performs exactly
those insertions on a COPY of the reference (never on /root/reference, never stored in git) and drops
our own sources (this directory) next to them.  Every edit is anchored on a short literal of the
reference and fails loudly if the anchor is missing, so a reference update cannot silently
produce a half-patched tree.  INTEGRATION.md shows the same edits as the patch a maintainer would apply.

usage: apply_overlay.py <scratch_reference_root> <repo_root>
"""
import shutil
import sys
from pathlib import Path


class Patch:
    def __init__(self, path: Path):
        self.path = path
        self.text = path.read_text()

    def insert_after(self, anchor: str, addition: str, occurrence: int = 0):
        idx = self._find(anchor, occurrence) + len(anchor)
        self.text = self.text[:idx] + addition + self.text[idx:]

    def insert_before(self, anchor: str, addition: str, occurrence: int = 0):
        idx = self._find(anchor, occurrence)
        self.text = self.text[:idx] + addition + self.text[idx:]

    def replace(self, anchor: str, new: str, occurrence: int = 0):
        idx = self._find(anchor, occurrence)
        self.text = self.text[:idx] + new + self.text[idx + len(anchor):]

    def _find(self, anchor: str, occurrence: int) -> int:
        idx = -1
        for _ in range(occurrence + 1):
            idx = self.text.find(anchor, idx + 1)
            if idx < 0:
                raise SystemExit(f"apply_overlay: anchor {anchor!r} (occurrence {occurrence}) not found in {self.path}")
        return idx

    def save(self):
        self.path.write_text(self.text)


def main():
    ref = Path(sys.argv[1])
    repo = Path(sys.argv[2])
    tf = ref / "TensorFrost"
    overlay = repo / "tensorfrost_b200" / "overlay"

    # 1. our sources: backend glue + emitter (TensorFrost/CMakeLists.txt globs *.cpp recursively)
    for rel in ["Backend/Backends/CUDA/CUDA.h", "Backend/Backends/CUDA/CudaBackend.cpp", "Backend/Backends/CUDA/CudaLibrary.cpp",
                "Backend/Backends/CUDA/CudaPython.cpp", "Backend/CodeGen/Langs/CUDA.cpp"]:
        dst = tf / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(overlay / rel, dst)

    # 2. enums + include (Backend/Backend.h)
    p = Patch(tf / "Backend" / "Backend.h")
    p.insert_after('#include "Backends/OpenGL/OpenGL.h"\n', '#include "Backends/CUDA/CUDA.h"\n')
    p.insert_after("\tOpenGL,\n", "\tCUDA,\n")           # enum class BackendType
    p.insert_after("\tGLSL,\n", "\tCUDA,\n")             # enum class CodeGenLang
    p.save()

    # 3. backend switch, kernel compilation, regions (Backend/Backend.cpp)
    p = Patch(tf / "Backend" / "Backend.cpp")
    p.insert_after("\t\t\t\tStopOpenGL();\n\t\t\t\tbreak;\n",
                   "\t\t\tcase BackendType::CUDA:\n\t\t\t\tStopCUDA();\n\t\t\t\tbreak;\n")
    p.insert_before("\t\tcase BackendType::OpenGL:\n\t\t\tStartOpenGL();",
                    "\t\tcase BackendType::CUDA:\n"
                    "\t\t\tStartCUDA();\n"
                    "\t\t\tcurrent_kernel_lang = CodeGenLang::CUDA;\n"
                    "\t\t\tcudaKernelCompileOptions = compilerOptions;  // NVRTC flags; the host program needs none\n"
                    "\t\t\tkernelCompileOptions = \"-O1\";\n"
                    "\t\t\tglobal_memory_manager = new CudaMemoryManager();\n"
                    "\t\t\tglobal_kernel_manager = new CudaKernelManager();\n"
                    "\t\t\tbreak;\n")
    # library lowerings follow the kernel LANGUAGE, so codegen mode with kernel_lang=tf.cuda_lang shows the same host program
    p.insert_after("\tif (kernelType != CodeGenLang::None) {\n\t\tcurrent_kernel_lang = kernelType;\n\t}\n",
                   "\tif (current_kernel_lang == CodeGenLang::CUDA) InstallCudaLibraryLowerings();\n")
    p.insert_after("auto start_time = chrono::high_resolution_clock::now();\n",
                   "\tif (current_backend == BackendType::CUDA) {\n"
                   "\t\t((CudaKernelManager*)global_kernel_manager)->CompileProgram(program);\n"
                   "\t}\n")
    p.insert_after("\t\t\tcase BackendType::CPU:\n\t\t\t\t//already in the host program\n\t\t\t\tbreak;\n",
                   "\t\t\tcase BackendType::CUDA:\n\t\t\t\t//compiled above, all kernels of the program at once\n\t\t\t\tbreak;\n")
    # every program execution is bracketed for the launch recorder (graph replay of the program's dispatch chain, include/tfcuda.h)
    p.insert_before("\ttry {\n\t\tprogram->execute_callback(", "\tCudaProgramScope cuda_scope;\n")
    p.insert_after("throw std::runtime_error(\"Error executing program \" + program->program_name + \": \" + e.what());\n\t}\n",
                   "\tcuda_scope.Finish();\n")
    p.insert_before("\tif (current_backend == BackendType::OpenGL) {\n\t\tif (begin) {",
                    "\tif (current_backend == BackendType::CUDA) {\n\t\tCudaRegion(name, begin);\n\t}\n")
    p.save()

    # 3b. host program: per-process file names + content-addressed cache instead of the fixed /tmp/generated_lib_<id>.cpp
    # (Backends/CPU/KernelCompiler.cpp:93-113); the reference path is untouched for every other backend
    p = Patch(tf / "Backend" / "Backends" / "CPU" / "KernelCompiler.cpp")
    p.insert_after('#include "KernelCompiler.h"\n', '#include "Backend/Backends/CUDA/CUDA.h"\n')
    p.insert_after("char* dllName, size_t program_id) {\n",
                   "\tif (CudaHostProgramCache(sourceCode, dllName, program_id)) return;\n")
    p.save()

    # 4. emitter dispatch (Backend/CodeGen/Generators.{h,cpp})
    p = Patch(tf / "Backend" / "CodeGen" / "Generators.cpp")
    p.insert_before("\t\tdefault:\n\t\t\tthrow std::runtime_error(\"Code generation for this language",
                    "\t\tcase CodeGenLang::CUDA:\n\t\t\tGenerateCUDAKernel(program, kernel);\n\t\t\treturn;\n")
    p.save()
    p = Patch(tf / "Backend" / "CodeGen" / "Generators.h")
    p.insert_after("void GenerateGLSLKernel(Program* program, Kernel* kernel);\n",
                   "void GenerateCUDAKernel(Program* program, Kernel* kernel);\n")
    p.save()

    # 4b. reductions the CUDA library takes whole are not split into the staged two-pass form (Steps/Optimization.cpp:471-510)
    p = Patch(tf / "Compiler" / "Steps" / "Optimization.cpp")
    p.insert_before("void IR::OptimizeReductions() {", "bool CudaLibraryWantsReduction(Node* node);  // Backend/Backends/CUDA/CudaLibrary.cpp\n\n")
    p.insert_after("\t\tint axis = (int)node->data[0];\n", "\t\tif (CudaLibraryWantsReduction(node)) continue;\n")
    p.save()

    # 4c. default thread-block shapes for CUDA (Steps/GraphOps.cpp:1143-1166)
    p = Patch(tf / "Compiler" / "Steps" / "GraphOps.cpp")
    p.insert_before("Tensor* IR::LinearBlockModeIndices(", "vector<int> CudaDefaultGroupSize(int dims, const vector<int>& const_shape, Node* kernel_node);  // Backend/CodeGen/Langs/CUDA.cpp\n\n")
    p.insert_before("\t\t\t//if the dimensions are known, then use the minimum of the group size and the shape",
                    "\t\t\t{\n\t\t\t\tvector<int> const_shape;\n"
                    "\t\t\t\tfor (int i = 0; i < dims; i++) const_shape.push_back(kernel_shape[i]->TryGetConstant());\n"
                    "\t\t\t\tvector<int> cuda_group = CudaDefaultGroupSize(dims, const_shape, kernel_);\n"
                    "\t\t\t\tif (!cuda_group.empty()) kernel_->group_size = cuda_group;\n\t\t\t}\n")
    p.save()

    # 5. python bindings (Frontend/Python/PybindModule.cpp)
    p = Patch(tf / "Frontend" / "Python" / "PybindModule.cpp")
    p.insert_after("void ModuleDefinitions(py::module& m);\n", "void CudaDefinitions(py::module& m);\n")
    p.insert_after('backend_type.value("opengl", BackendType::OpenGL);\n', '\tbackend_type.value("cuda", BackendType::CUDA);\n')
    p.insert_after('code_gen_lang.value("hlsl", CodeGenLang::HLSL);\n', '\tcode_gen_lang.value("cuda", CodeGenLang::CUDA);\n')
    p.insert_after('m.attr("opengl") = BackendType::OpenGL;\n', '\tm.attr("cuda") = BackendType::CUDA;\n')
    p.insert_after('m.attr("hlsl_lang") = CodeGenLang::HLSL;\n', '\tm.attr("cuda_lang") = CodeGenLang::CUDA;\n')
    p.insert_after("\tModuleDefinitions(m);\n", "\tCudaDefinitions(m);\n")
    p.save()

    # 5b. bulk paths behind tf.tensor(np) and .numpy (Frontend/Python/PyTensorMemory.{cpp,h}): contiguous 4-byte arrays skip the
    # per-element conversion; everything else falls through to the reference code
    p = Patch(tf / "Frontend" / "Python" / "PyTensorMemory.cpp")
    p.insert_before("PyTensorMemory::PyTensorMemory(py::array arr) {", "bool CudaBulkTensor(const py::buffer_info& info, TFTensor** out);  // Backend/Backends/CUDA/CudaPython.cpp\n\n")
    p.insert_after("PyTensorMemory::PyTensorMemory(py::array arr) {\n    py::buffer_info info = arr.request();\n",
                   "    if (CudaBulkTensor(info, &tensor_)) return;\n")
    p.save()
    p = Patch(tf / "Frontend" / "Python" / "PyTensorMemory.h")
    p.insert_before("// Tensor wrapper for python\nclass PyTensorMemory {", "bool CudaBulkReadback(const TFTensor* t, void* dst, size_t item_size);  // Backend/Backends/CUDA/CudaPython.cpp\n\n")
    p.insert_after("\t\tpy::array_t<T> arr(shape);\n", "\t\tif (CudaBulkReadback(tensor_, arr.request().ptr, sizeof(T))) return arr;\n")
    p.save()

    # 6. build: link libtfcuda.so (found next to the module at run time) and see include/tfcuda.h
    p = Patch(tf / "CMakeLists.txt")
    p.text += (
        "\n# --- tensorfrost_b200 overlay ---\n"
        f"target_include_directories(TensorFrost PRIVATE {repo / 'include'})\n"
        f"target_link_libraries(TensorFrost PRIVATE {repo / 'tensorfrost_b200' / 'lib' / 'libtfcuda.so'})\n"
        "set_target_properties(TensorFrost PROPERTIES BUILD_WITH_INSTALL_RPATH ON INSTALL_RPATH \"$ORIGIN\")\n"
    )
    # glad's generator otherwise downloads gl.xml; REPRODUCIBLE uses the vendored spec
    p.replace("glad_gl_core_46 SHARED API", "glad_gl_core_46 STATIC REPRODUCIBLE API")
    p.save()

    print("apply_overlay: patched", ref)


if __name__ == "__main__":
    main()
