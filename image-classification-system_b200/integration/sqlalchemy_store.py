"""``ImageStore`` on the reference's own SQLAlchemy session and models — the adapter INTEGRATION.md installs.

    from app.db.models import Imagem, ConjuntoImagens                      # the reference's models, unchanged
    from ics_b200.integration.sqlalchemy_store import SqlAlchemyImageStore
    store = SqlAlchemyImageStore(db, Imagem, ConjuntoImagens)              # db = the request / thread session
    WebDAVSync(nextcloud_client, store)._process_image_batch(images, folder_path, conjunto_id)

The storage engine itself is out of scope; this file only maps the seven operations of the seam
onto the session calls the reference makes at the same places:

  get / get_many    db.query(Imagem).filter_by(content_hash=...).first()          webdav_sync.py:324
                    db.query(Imagem).filter(Imagem.content_hash.in_(...)).all()   (ONE lookup per batch instead of n)
  insert            db.add(Imagem(**row)); db.flush()                             webdav_sync.py:352-353
                    IntegrityError on flush -> store.DuplicateKeyError            webdav_sync.py:355 (the caller then
                    calls rollback() and merges, exactly as the reference does)
  update            attribute assignment on the loaded row                       webdav_sync.py:373-398
                    (+ flag_modified for the JSONB column when it was mutated in place)
  commit / rollback db.commit() / db.rollback()                                   webdav_sync.py:282, :357, :423
  folder_for        lookup-or-create ConjuntoImagens by file_id                   activity_api_sync.py:818-841

Rows cross the seam as plain dicts with the ``Imagem`` column names (app/db/models.py:202-222); the dict handed out by
``get`` shares its ``metadados`` object with the ORM row, which is what lets the services update it in place like the
reference does.  SQLAlchemy is imported lazily and only for ``IntegrityError`` / ``flag_modified``: the module loads
(and is tested) against any session object with the same five methods.
"""
from __future__ import annotations

import uuid
from typing import Dict, Iterable, Optional

from ..store import DuplicateKeyError

IMAGEM_COLUMNS = ("content_hash", "nome_img", "caminho_img", "metadados", "existe_no_nextcloud", "data_proc",
                  "data_sinc", "id_cnj")


def _is_integrity_error(e: BaseException) -> bool:
    try:
        from sqlalchemy.exc import IntegrityError
        if isinstance(e, IntegrityError):
            return True
    except ImportError:
        pass
    return type(e).__name__ == "IntegrityError"             # any DB-API / stub class of that name


def _flag_modified(obj, name: str) -> None:
    try:
        from sqlalchemy.orm.attributes import flag_modified
        flag_modified(obj, name)                             # JSONB mutated in place: tell the unit of work
    except Exception:  # noqa: BLE001 - not a mapped instance (tests' stub models): plain attribute, nothing to flag
        pass


class SqlAlchemyImageStore:
    def __init__(self, session, imagem_model, conjunto_model=None):
        self.db, self.Imagem, self.Conjunto = session, imagem_model, conjunto_model

    # ---- rows <-> dicts
    @staticmethod
    def _row(obj) -> Dict:
        return {c: getattr(obj, c, None) for c in IMAGEM_COLUMNS}

    def _load(self, content_hash: str):
        return self.db.query(self.Imagem).filter_by(content_hash=content_hash).first()

    # ---- ImageStore
    def get(self, content_hash: str) -> Optional[Dict]:
        obj = self._load(content_hash)
        return None if obj is None else self._row(obj)

    def get_many(self, content_hashes: Iterable[str]) -> Dict[str, Dict]:
        keys = list(set(content_hashes))
        if not keys:
            return {}
        rows = self.db.query(self.Imagem).filter(self.Imagem.content_hash.in_(keys)).all()
        return {r.content_hash: self._row(r) for r in rows}

    def insert(self, row: Dict) -> None:
        self.db.add(self.Imagem(**row))
        try:
            self.db.flush()                                  # surfaces a concurrent insert before the commit (:353)
        except Exception as e:  # noqa: BLE001
            if _is_integrity_error(e):
                raise DuplicateKeyError(f"duplicate primary key {row['content_hash']}") from e
            raise

    def update(self, content_hash: str, fields: Dict) -> None:
        obj = self._load(content_hash)
        if obj is None:
            raise KeyError(content_hash)
        for k, v in fields.items():
            setattr(obj, k, v)
        if "metadados" in fields:
            _flag_modified(obj, "metadados")

    def commit(self) -> None:
        self.db.commit()

    def rollback(self) -> None:
        self.db.rollback()

    def folder_for(self, file_id: str, name: str, path: str, now) -> Optional[Dict]:
        if self.Conjunto is None:
            return None
        obj = self.db.query(self.Conjunto).filter_by(file_id=file_id).first()
        if obj is None:
            obj = self.Conjunto(id_cnj=uuid.uuid4(), nome_conj=name, caminho_conj=path, file_id=file_id,
                                imagens_sincronizadas=False, existe_no_nextcloud=True, data_proc=now, data_sinc=now)
            self.db.add(obj)
            self.db.flush()
        return {"id_cnj": obj.id_cnj, "nome_conj": obj.nome_conj, "caminho_conj": obj.caminho_conj, "file_id": obj.file_id}
