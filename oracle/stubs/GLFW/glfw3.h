/* oracle/stubs/GLFW/glfw3.h -- TEST INFRASTRUCTURE ONLY; see ../GL/gl.h.  Window/context calls of
 * gpgpu_gl4.h:142-174 (coInit / coTerm) compile to nothing; the shim never calls them. */
#ifndef UGEMM_ORACLE_STUB_GLFW_H
#define UGEMM_ORACLE_STUB_GLFW_H
typedef struct GLFWwindow GLFWwindow;
enum {
	GLFW_FALSE = 0, GLFW_TRUE = 1, GLFW_CONTEXT_VERSION_MAJOR, GLFW_CONTEXT_VERSION_MINOR, GLFW_OPENGL_FORWARD_COMPAT,
	GLFW_OPENGL_PROFILE, GLFW_OPENGL_CORE_PROFILE, GLFW_VISIBLE
};
#define glfwInit(...)               0
#define glfwWindowHint(...)         ((void)0)
#define glfwCreateWindow(...)       ((GLFWwindow *)0)
#define glfwTerminate(...)          ((void)0)
#define glfwMakeContextCurrent(...) ((void)0)
#endif
