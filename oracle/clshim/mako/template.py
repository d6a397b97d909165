"""ORACLE — TEST INFRASTRUCTURE ONLY.  `mako.template.Template(filename=...).render(**kw)` for templates that use
nothing but `${name}` substitutions -- which is all kernel_farfield.cl / kernel_nearfield.cl contain
(`${my_dtype}`, `${f_native}`; calc.py:605-624).  Any other Mako syntax raises."""
import re


class Template:
    def __init__(self, text=None, filename=None):
        if filename is not None:
            with open(filename) as f:
                text = f.read()
        self.text = text
        for marker in ('<%', '%>', '##'):
            if marker in text:
                raise NotImplementedError(f'clshim mako: unsupported template syntax {marker!r}')
        if re.search(r'^\s*%', text, re.M):
            raise NotImplementedError('clshim mako: control lines are not supported')

    def render(self, **kw):
        def sub(m):
            return str(kw[m.group(1).strip()])
        return re.sub(r'\$\{([^}]*)\}', sub, self.text)
