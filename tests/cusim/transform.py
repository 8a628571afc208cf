"""TEST INFRASTRUCTURE - rewrites the CUDA-only syntax of a .cu file so that g++ can compile it against tests/cusim/cuda_runtime.h:

    kernel<T...><<<grid, block, smem, stream>>>(args)   ->   cusim::launch(grid, block, smem, stream, [=] { kernel<T...>(args); })
    extern __shared__ [__align__(n)] T name[];          ->   T* name = reinterpret_cast<T*>(cusim::dyn_smem);

Everything else (kernels, device helpers, host launch wrappers) is compiled as written."""
import re


def _balanced(src, i, open_ch='(', close_ch=')'):
    """index just past the bracket group that opens at src[i]"""
    assert src[i] == open_ch, src[i:i + 20]
    depth = 0
    while True:
        ch = src[i]
        depth += ch == open_ch
        depth -= ch == close_ch
        i += 1
        if depth == 0:
            return i


def transform(src):
    src = re.sub(r'extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+(?:\s+\w+)?)\s+(\w+)\[\];', r'\1* \2 = reinterpret_cast<\1*>(cusim::dyn_smem);', src)
    out, pos = [], 0
    while True:
        i = src.find('<<<', pos)
        if i < 0:
            out.append(src[pos:])
            return ''.join(out)
        j = src.index('>>>', i)
        cfg = src[i + 3:j]
        # kernel name (+ template arguments) immediately before <<<
        k = i
        if src[k - 1] == '>':
            depth, k = 0, k - 1
            while True:
                depth += src[k] == '>'
                depth -= src[k] == '<'
                if depth == 0:
                    break
                k -= 1
        m = re.search(r'[A-Za-z_]\w*$', src[:k])
        name_start = m.start()
        name = src[name_start:i]
        a0 = j + 3
        a1 = _balanced(src, a0)
        args = src[a0 + 1:a1 - 1]
        out.append(src[pos:name_start])
        out.append(f'cusim::launch({cfg}, [=] {{ {name}({args}); }})')
        pos = a1
