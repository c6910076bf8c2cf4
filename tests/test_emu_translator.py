"""The launch-syntax rewriter of tests/emu/build.py (pure Python, no compiler needed)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import build as emu_build  # noqa: E402


def test_plain_launch_with_all_four_config_arguments():
    out = emu_build.translate("    pack_kernel<<<blocks_for(cover), THREADS, 0, h->stream>>>(h->d_state, d_sigma2, h->m);\n")
    assert out == ('    emu::launch("pack_kernel", dim3(blocks_for(cover)), dim3(THREADS), (size_t)(0), (cudaStream_t)(h->stream), '
                   '[&]() { pack_kernel(h->d_state, d_sigma2, h->m); });\n')


def test_template_arguments_and_short_config():
    out = emu_build.translate("if (c) pass1_kernel<true, false><<<g, 256>>>(a, (int)n, f(x, y));")
    assert 'emu::launch("pass1_kernel<true, false>", dim3(g), dim3(256), (size_t)(0), (cudaStream_t)(nullptr)' in out
    assert "[&]() { pass1_kernel<true, false>(a, (int)n, f(x, y)); })" in out and out.startswith("if (c) ")


def test_parenthesised_kernel_expression_and_multi_line_arguments():
    src = "KERNEL_OF((probe<11, false>))\n(probe<11, false>)<<<dim3(a, b), THREADS, smem(1, 2), s>>>(d,\n        500, 1.0f);\nnext();"
    out = emu_build.translate(src)
    assert "dim3(dim3(a, b)), dim3(THREADS), (size_t)(smem(1, 2)), (cudaStream_t)(s)" in out
    assert "{ (probe<11, false>)(d,\n        500, 1.0f); })" in out and out.endswith(";\nnext();")


def test_text_without_launches_is_untouched_and_every_product_launch_is_rewritten():
    assert emu_build.translate("a << b; c >> d; x <<= 3;") == "a << b; c >> d; x <<= 3;"
    csrc = os.path.join(emu_build.ROOT, "probreg_b200", "csrc")
    for name in os.listdir(csrc):
        if name.endswith((".cu", ".inl")):
            text = open(os.path.join(csrc, name)).read()
            out = emu_build.translate(text)
            assert "<<<" not in out and ">>>" not in out.replace(">>>=", ""), name
            assert out.count("emu::launch(") == text.count("<<<"), name
