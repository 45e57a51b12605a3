"""Build numpy structured dtypes (C layout, natural alignment) from the POD
struct declarations in a C header, so that the Python side never hand-mirrors
`msv_config` / `msv_env_state` (include/masurv.h).  The library exports
`msv_sizeof_*` so the result is checked against the compiler's layout."""
import re
import numpy as np

_PRIM = {
    'int32_t': np.int32, 'uint32_t': np.uint32, 'int64_t': np.int64,
    'uint64_t': np.uint64, 'uint8_t': np.uint8, 'int8_t': np.int8,
    'float': np.float32, 'double': np.float64, 'int': np.int32,
}


def _strip_comments(text):
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    return re.sub(r'//[^\n]*', ' ', text)


def parse_header(*paths):
    """Return (defines: dict[str,int], structs: dict[str,np.dtype])."""
    text = '\n'.join(_strip_comments(open(p).read()) for p in paths)
    defines = {}
    for m in re.finditer(r'^[ \t]*#define[ \t]+(\w+)[ \t]+(.+?)[ \t]*$', text, flags=re.M):
        name, expr = m.group(1), m.group(2)
        if not re.fullmatch(r'[\w\s()+\-*/<>]+', expr):
            continue
        try:
            val = eval(expr.replace('/', '//'), {'__builtins__': {}}, dict(defines))
        except Exception:
            continue
        if isinstance(val, int):
            defines[name] = val
    structs = {}
    for m in re.finditer(r'typedef\s+struct\s+\w*\s*\{(.*?)\}\s*(\w+)\s*;', text, flags=re.S):
        body, name = m.group(1), m.group(2)
        if '*' in re.sub(r'\[[^\]]*\]', '', body):
            continue  # structs holding pointers are mirrored by hand
        fields = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            dm = re.match(r'(const\s+)?(\w+)\s+(.*)$', decl, flags=re.S)
            tname, rest = dm.group(2), dm.group(3)
            if tname in _PRIM:
                base = np.dtype(_PRIM[tname])
            elif tname in structs:
                base = structs[tname]
            else:
                raise ValueError(f'unknown type {tname} in struct {name}')
            for var in rest.split(','):
                var = var.strip()
                vm = re.match(r'(\w+)((?:\s*\[[^\]]+\])*)$', var)
                fname = vm.group(1)
                dims = tuple(
                    int(eval(d.replace('/', '//'), {'__builtins__': {}}, dict(defines)))
                    for d in re.findall(r'\[([^\]]+)\]', vm.group(2)))
                fields.append((fname, base, dims) if dims else (fname, base))
        structs[name] = np.dtype(fields, align=True)
    return defines, structs
