/* oracle/stubs/GL/gl.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A do-nothing stand-in for <GL/gl.h> so that the reference's OpenGL-compute header
 * (sgemm_gl1.h -> gpgpu_gl4.h:29-31) compiles UNMODIFIED in an image without Mesa / GLFW.  Only the plain-C
 * CPU code of that header is ever executed through oracle/ref_conv_shim.c (im2col, sgemm_gl1.h:166-190);
 * every GL entry point is a macro that evaluates to 0 and the shim never calls a function that uses one.
 */
#ifndef UGEMM_ORACLE_STUB_GL_H
#define UGEMM_ORACLE_STUB_GL_H
typedef unsigned int GLuint;
typedef int GLint;
typedef int GLsizei;
typedef unsigned int GLenum;
typedef void GLvoid;
typedef unsigned char GLubyte;
enum {
	GL_FALSE = 0, GL_TRUE = 1, GL_COMPUTE_SHADER, GL_COMPILE_STATUS, GL_INFO_LOG_LENGTH, GL_LINK_STATUS,
	GL_SHADER_STORAGE_BUFFER, GL_DYNAMIC_COPY, GL_WRITE_ONLY, GL_RENDERER, GL_VERSION, GL_CLAMP_TO_BORDER
};
#define glCreateShader(...)       0u
#define glShaderSource(...)       ((void)0)
#define glCompileShader(...)      ((void)0)
#define glGetShaderiv(...)        ((void)0)
#define glGetShaderInfoLog(...)   ((void)0)
#define glDeleteShader(...)       ((void)0)
#define glCreateProgram(...)      0u
#define glAttachShader(...)       ((void)0)
#define glLinkProgram(...)        ((void)0)
#define glGetProgramiv(...)       ((void)0)
#define glGetProgramInfoLog(...)  ((void)0)
#define glDeleteProgram(...)      ((void)0)
#define glDetachShader(...)       ((void)0)
#define glUseProgram(...)         ((void)0)
#define glGetUniformLocation(...) 0
#define glUniform1iv(...)         ((void)0)
#define glDispatchCompute(...)    ((void)0)
#define glGenBuffers(...)         ((void)0)
#define glDeleteBuffers(...)      ((void)0)
#define glBindBuffer(...)         ((void)0)
#define glBindBufferBase(...)     ((void)0)
#define glBufferData(...)         ((void)0)
#define glBufferSubData(...)      ((void)0)
#define glGetBufferSubData(...)   ((void)0)
#define glMapBuffer(...)          ((void *)0)
#define glUnmapBuffer(...)        ((void)0)
#define glMemoryBarrier(...)      ((void)0)
#define glFinish(...)             ((void)0)
#define glGetString(...)          ((const GLubyte *)"stub")
#endif
