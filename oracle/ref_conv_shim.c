/* oracle/ref_conv_shim.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Export shim around the reference's UNMODIFIED convolution host code: sgemm_gl1.h is compiled where it lies
 * (-I$UGEMM_REF) against the do-nothing GL/GLFW stand-ins of oracle/stubs/ (the image has no Mesa / GLFW headers),
 * so its plain-C `im2col` (sgemm_gl1.h:166-190) runs on the CPU exactly as shipped.  The GEMM in
 * gl_convolution_LReLU (sgemm_gl1.h:192-218) is an OpenGL compute dispatch and cannot run here; ref_gl_convolution
 * therefore chains the reference's own pieces the way that function does -- reference im2col -> reference CPU SGEMM
 * (sgemm_sse / sgemm_cpu from ugemm.h, M = ch, N = hcol*wcol, K = k*k*ich, sgemm_gl1.h:200-207) -> the bias +
 * LeakyReLU(0.1) loop restated from sgemm_gl1.h:210-217 -- and pins oracle_im2col / oracle_convolution.
 *
 * Built into its own shared object (oracle/_ref/libugemm_ref_conv.so) because sgemm_gl1.h:191 carries a 2 GiB
 * `workspace` global (zero-fill BSS, never touched here) that the main reference library should not drag along.
 */
#include <stdio.h>
#include <stdlib.h>
#include "ugemm.h"
#include "sgemm_gl1.h"

void ref_im2col(const float *im, int channels, int height, int width, int kernel_h, int kernel_w,
                int pad_h, int pad_w, int stride_h, int stride_w, float *col)
{ im2col(im, channels, height, width, kernel_h, kernel_w, pad_h, pad_w, stride_h, stride_w, col); }

/* which: 0 = sgemm_cpu (naive), 1 = sgemm_sse.  bias == NULL: plain convolution (ocl_convolution, sgemm_ocl1.h:271-300);
 * otherwise the bias + LeakyReLU(0.1) tail of gl_convolution_LReLU (sgemm_gl1.h:210-217). */
int ref_gl_convolution(int which, const float *inputs, int ich, int w, int h, const float *weights, int k, int pad, int stride,
                       float *outputs, int ch, const float *bias)
{
	int hcol = (h + 2 * pad - k) / stride + 1;
	int wcol = (w + 2 * pad - k) / stride + 1;
	int M = ch, N = wcol * hcol, K = k * k * ich;
	float *col = (float *)malloc(sizeof(float) * (size_t)K * N);
	if (!col) return 1;
	im2col(inputs, ich, h, w, k, k, pad, pad, stride, stride, col);
	/* sgemm_cpu computes 0*C for beta == 0 (ugemm.h:313) and so propagates whatever the caller left in outputs: start from zeros */
	if (which == 0) for (int i = 0; i < M * N; i++) outputs[i] = 0.f;
	if (which == 0) sgemm_cpu('R', 'N', 'N', M, N, K, 1.f, weights, K, col, N, 0.f, outputs, N);
	else            sgemm_sse('R', 'N', 'N', M, N, K, 1.f, weights, K, col, N, 0.f, outputs, N);
	free(col);
	if (bias) {
		float *p = outputs;
		for (int i = 0; i < ch; i++)
			for (int n = 0; n < N; n++) { *p += bias[i]; *p = *p > 0 ? (*p) : (*p) * 0.1; p++; }
	}
	return 0;
}
