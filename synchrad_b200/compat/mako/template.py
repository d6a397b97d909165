"""`mako.template.Template(filename=...).render(**kw)` for templates made of `${name}` substitutions only, which is
all the reference's kernel files contain (calc.py:605-624).  Only needed where Mako itself is not installed; the
rendered source is merely inspected by compat/pyopencl (kernel names, compute type)."""
import re


class Template:
    def __init__(self, text=None, filename=None):
        if filename is not None:
            with open(filename) as f:
                text = f.read()
        for marker in ('<%', '%>'):
            if marker in text:
                raise NotImplementedError(f'unsupported template syntax {marker!r}: install Mako')
        self.text = text

    def render(self, **kw):
        return re.sub(r'\$\{([^}]*)\}', lambda m: str(kw[m.group(1).strip()]), self.text)
